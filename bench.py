#!/usr/bin/env python
"""bench.py — CG iterations/sec and HBM roofline of the liblcg hot path on B200 (driver contract in README/DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--iters I]

A "step" is ONE SOLVE of a fixed number of iterations (`--iters`, default 200: epsilon = 1e-300 and
max_iterations = I, so every step runs exactly I iterations of the reference's loop and returns
LCG_REACHED_MAX_ITERATIONS) of the workload's system.  Workloads (BASELINE.json configs):

    pcg27_256      27-point Poisson 256^3 (16.8 M rows, 449 M nnz), Jacobi-PCG      [default; configs[3], the metric's config]
    cg7_128        7-point Poisson 128^3 (2.1 M rows), CG                           [configs[2]]
    bicgstab7cd_512  7-point convection-diffusion 512^3 (134 M rows), BiCGSTAB      [configs[4]]
    cgs7cd_512     same system, CGS

value   = iterations/s with the system, m and B resident in HBM (handle-shaped lcgb200_solve, device vectors).
e2e     = iterations/s through the reference-shaped entry point (lcg_solver_preconditioned_cuda & co. with the
          sentinel callbacks): HOST m/B in, HOST m out, the host<->device copies inside the timed region.
roofline = the SpMV(+fused dot) kernel: algorithmic bytes per launch / its CUDA-event duration, measured in a
          second pass of the same steps with lcgb200_set_profile(1).
cpu_baseline = the UNMODIFIED reference CPU/OpenMP solver (oracle/_ref, kind "reference"; the C port if that
          library is absent) on the same system for a bounded number of iterations, all host threads.

With --impl reference only that CPU solver runs (rank 0 only under torchrun).
N > 1: one process per GPU (torchrun), the matrix row-partitioned in z-slabs, halo exchange + scalar allreduce
over NCCL; total work fixed -> "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (stencil kind, grid, solver name)
    "pcg27_256": ("27pt", 256, "PCG"),
    "cg7_128": ("7pt", 128, "CG"),
    "bicgstab7cd_512": ("7pt_cd", 512, "BICGSTAB"),
    "cgs7cd_512": ("7pt_cd", 512, "CGS"),
    # the reference's own sample systems (configs[0], configs[1]): launch-latency-bound, it/s only
    "case10k_cg": ("fixture:10K", 0, "CG"),
    "case10k_pcg": ("fixture:10K", 0, "PCG"),
    # small variants for quick checks (not bench lines)
    "pcg27_64": ("27pt", 64, "PCG"),
    "pcg27_128": ("27pt", 128, "PCG"),   # the per-GPU share of pcg27_256 on 8 GPUs (2.1 M rows), without the exchanges
    "pcg27_160": ("27pt", 160, "PCG"),   # ~ the share on 4 GPUs
    "pcg27_200": ("27pt", 200, "PCG"),   # ~ the share on 2 GPUs
    "bicgstab7cd_96": ("7pt_cd", 96, "BICGSTAB"),
}
KIND_ID = {"7pt": 0, "27pt": 1, "7pt_cd": 2}
SOLVER_ID = {"CG": 0, "PCG": 1, "CGS": 2, "BICGSTAB": 3, "BICGSTAB2": 4}


def stencil_nnz(kind, g):
    return (3 * g - 2) ** 3 if kind == "27pt" else 7 * g ** 3 - 6 * g ** 2


def bytes_per_iteration(solver, n, nnz):
    """Algorithmic HBM bytes of one iteration (double + int32 CSR)."""
    spmv = 12 * nnz + 4 * (n + 1) + 16 * n
    extra = {"CG": (1, 9), "PCG": (1, 11), "CGS": (2, 17), "BICGSTAB": (2, 16), "BICGSTAB2": (2, 16)}[solver]
    return extra[0] * spmv + extra[1] * 8 * n


def spmv_bytes(n, nnz):
    return 12 * nnz + 4 * (n + 1) + 16 * n


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


# ------------------------------------------------------------------------------------ CPU arms (reference)
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_system(kind, g):
    """The workload's system on the host: a SURVEY 8(d) stencil or one of the reference's fixtures (tests/golden/data)."""
    if kind.startswith("fixture:"):
        from liblcg_b200 import io as lio
        return lio.load_fixture(kind.split(":")[1])
    from oracle import pyoracle as po
    return po.gen_system(kind, g)


def cpu_reference_run(kind, g, solver, iters, repeats=1, warmup=0):
    """Times the reference's own CPU/OpenMP solver (oracle/_ref) — or the C port when that library is absent — on
    the workload's system for `iters` iterations per solve.  Returns (it_per_s list, kind, cores)."""
    from oracle import pyoracle as po
    which = "reference" if po.have_reference() else "port"
    orc = po.Oracle(which)
    # all host cores this process may use, whatever OMP_NUM_THREADS a launcher exported (torchrun sets it to 1)
    orc.set_num_threads(host_cores())
    S = host_system(kind, g)
    diag = None
    if solver == "PCG":
        from liblcg_b200 import io as lio
        diag = lio.csr_diagonal(S["row_ptr"], S["col"], S["val"]) if kind.startswith("fixture:") else np.full(S["n"], 26.0 if kind == "27pt" else 6.0)
    para = po.default_para(epsilon=1e-300, max_iterations=iters)
    rates = []
    for i in range(warmup + repeats):
        r = orc.solve(SOLVER_ID[solver], S, S["b"], para=para, diag=diag, progress=False)
        assert r.ret == -1019, r.ret
        if i >= warmup:
            rates.append(iters / r.seconds)
    return rates, which, orc.num_threads()


def run_reference_arm(args, wl):
    kind, g, solver = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if kind.startswith("fixture:"):
        S0 = host_system(kind, g)
        n, nnz = S0["n"], S0["nnz"]
    else:
        n, nnz = g ** 3, stencil_nnz(kind, g)
    # bounded sample: a few iterations of the same system per step, so K + W steps end within minutes
    it = args.ref_iters
    t0 = time.time()
    rates, which, cores = cpu_reference_run(kind, g, solver, it, repeats=args.steps, warmup=args.warmup)
    total_it = it * len(rates)
    secs = sum(it / r for r in rates)
    value = total_it / secs
    sample = f"{it} {solver} iterations per step of the same {kind} {g}^3 system (max_iterations={it}, epsilon=1e-300), {which} CPU/OpenMP solver, OpenMP CSR Ax callback"
    line = {
        "impl": "reference", "metric": "cg_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(rates), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "stencil": kind, "grid": g, "rows": n, "nnz": nnz, "solver": solver,
                   "iterations_per_step": it, "note": "CPU arm: whole system in host memory, larger than any cache"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": which, "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Process-wide state of the GPU arm: rank, device, torch.distributed handle."""
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (liblcg_b200 has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, prepare=None):
        """K steps, each bracketed by barrier + synchronize and a CUDA-event pair on the launching stream; the step times
        are summed and the MAX over ranks is taken.  `prepare` (resetting the initial guess) runs between the brackets."""
        torch = self.torch
        total, res = 0.0, []
        for _ in range(steps):
            if prepare is not None:
                prepare()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res.append(fn())
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        if self.dist is not None:
            t = torch.tensor([total], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            total = float(t.item())
        self.barrier()
        return total, res


def device_system(ctx, kind, g, solver, compress=False, keep_csr=False):
    """The workload's system generated on the device: the whole matrix (1 GPU) or this rank's z-slab of rows.
    Returns dict(op, b, n, nnz, n_loc, part, transport, dev_csr)."""
    import ctypes as C
    from liblcg_b200 import api, _lib
    torch, dev = ctx.torch, ctx.dev
    lib = _lib.load()
    n, nnz = g ** 3, stencil_nnz(kind, g)
    out = dict(n=n, nnz=nnz, part=None, transport=None, dev_csr=None)
    if ctx.world > 1:
        from liblcg_b200 import dist as ldist
        part = ldist.build_stencil_partition(kind, g, ctx.rank, ctx.world, dev, jacobi=(solver == "PCG"), compress=compress)
        out.update(op=part.op, b=part.b, n_loc=part.n_local, part=part, r0=part.plan.r0,
                   transport="nvlink-p2p (halo pushed into peer memory by the kernel that writes the SpMV input, reduction totals pushed from the kernel tails)"
                   if part.p2p else "nccl (send/recv halo + allreduce)")
        return out
    nz = C.c_longlong()
    assert lib.lcgb200_gen_stencil(KIND_ID[kind], g, 0, n, None, None, None, 0, C.byref(nz), None) == 0
    assert nz.value == nnz
    rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    assert lib.lcgb200_gen_stencil(KIND_ID[kind], g, 0, n, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), 0, None, None) == 0
    b_d = torch.empty(n, dtype=torch.float64, device=dev)
    assert lib.lcgb200_gen_rhs(KIND_ID[kind], g, 0, n, b_d.data_ptr(), None) == 0
    torch.cuda.synchronize()
    op = api.CsrOperator(rp, ci, va, jacobi=(solver == "PCG"), compress=compress)
    out.update(op=op, b=b_d, n_loc=n, r0=0, dev_csr=(rp, ci, va) if keep_csr else None)
    return out


def close_system(ctx, S):
    if S.get("part") is not None:
        S["part"].close()
    else:
        S["op"].close()
    S.clear()
    ctx.torch.cuda.empty_cache()


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def parity_block(ctx, kind, solver, g=48, k=25, compress=False):
    """Driver-visible correctness of THIS run's path (1 GPU or the row-partitioned handle over all ranks): exactly k
    iterations of the workload's solver on a g^3 system of the workload's stencil, against the CPU oracle port on rank 0.
    North-star bar: solution rel-L2 <= 1e-8, same iteration count, same return code."""
    from liblcg_b200 import api
    torch = ctx.torch
    sid = SOLVER_ID[solver]
    S = device_system(ctx, kind, g, solver, compress=compress)   # --compress: the twin runs on the compressed operator copy too
    fmt_level = S["op"].format()["level"]
    m_d = torch.zeros(S["n_loc"], dtype=torch.float64, device=ctx.dev)
    para = api.lcg_default_parameters(epsilon=1e-300, max_iterations=k)
    r = api.solve(S["op"], sid, m_d, S["b"], param=para, device=True, jacobi=(solver == "PCG"), stream=ctx.stream)
    torch.cuda.synchronize()
    x_loc = m_d.cpu().numpy()
    parts = [(S["r0"], x_loc)]
    if ctx.dist is not None:
        gathered = [None] * ctx.world
        ctx.dist.all_gather_object(gathered, (S["r0"], x_loc))
        parts = gathered
    err = api.last_error() if r.ret != api.LCG_REACHED_MAX_ITERATIONS else ""
    # one GPU: the same solve once more in reference-order mode (lcgb200_set_reference_order: no FMA contraction, serial row
    # sums and dot products) — iterate and final residual must equal the CPU port's bit for bit
    x_exact = r_exact = None
    if ctx.world == 1:
        m_d.zero_()
        api.set_reference_order(True)
        try:
            r_exact = api.solve(S["op"], sid, m_d, S["b"], param=para, device=True, jacobi=(solver == "PCG"), stream=ctx.stream)
        finally:
            api.set_reference_order(False)
        torch.cuda.synchronize()
        x_exact = m_d.cpu().numpy()
    close_system(ctx, S)
    if ctx.rank != 0:
        return None
    from oracle import pyoracle as po
    x = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])])
    H = po.gen_system(kind, g)
    diag = np.full(H["n"], 26.0 if kind == "27pt" else 6.0) if solver == "PCG" else None
    cpu = po.Oracle("port").solve(sid, H, H["b"], para=po.default_para(epsilon=1e-300, max_iterations=k), diag=diag)
    rel = float(np.linalg.norm(x - cpu.x) / max(np.linalg.norm(cpu.x), 1e-300))
    exact = None
    if x_exact is not None:
        exact = {"bit_identical_solution": bool(x_exact.tobytes() == cpu.x.tobytes()), "bit_identical_residual": bool(r_exact.residual == cpu.residual),
                 "ret": r_exact.ret, "iterations": r_exact.iterations,
                 "note": "lcgb200_set_reference_order(1): second build of the loops without FMA contraction and with serial sums"}
    return {"system": f"{kind} {g}^3", "solver": solver, "pinned_iterations": k, "n_gpus": ctx.world, "rel_l2": rel, "reference_order": exact,
            "operator_format": {0: "csr", 1: "dictionary codes", 2: "row patterns"}.get(fmt_level, str(fmt_level)),
            "ret_gpu": r.ret, "ret_cpu": cpu.ret, "iterations_gpu": r.iterations, "iterations_cpu": cpu.iters,
            "residual_gpu": r.residual, "residual_cpu": cpu.residual, "ok": bool(rel <= 1e-8 and r.ret == cpu.ret and r.iterations == cpu.iters),
            "checker": "oracle/lcg_oracle.c (CPU port, pinned bit-for-bit to the reference)", "error": err}


def extra_leg(ctx, name, steps, iters):
    """A further BASELINE config measured in the same run and printed inside the same JSON line: device-resident solves
    (the definition of `value`), the SpMV roofline from a profiled pass, and a parity block on a 48^3 twin of the system."""
    from liblcg_b200 import api
    torch = ctx.torch
    kind, g, solver = WORKLOADS[name]
    sid = SOLVER_ID[solver]
    S = device_system(ctx, kind, g, solver)
    n, nnz = S["n"], S["nnz"]
    op, b_d = S["op"], S["b"]
    m_d = torch.zeros(S["n_loc"], dtype=torch.float64, device=ctx.dev)
    para = api.lcg_default_parameters(epsilon=1e-300, max_iterations=iters)

    def step():
        r = api.solve(op, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=ctx.stream)
        if r.ret != api.LCG_REACHED_MAX_ITERATIONS or r.iterations != iters:
            raise RuntimeError(f"{name}: solve returned {r.ret} after {r.iterations} iterations: {api.last_error()}")
        return r

    for _ in range(3):
        m_d.zero_(); step()
    ms, res = ctx.timed(step, steps, prepare=lambda: m_d.zero_())
    value = steps * iters / (ms * 1e-3)
    api.set_profile(True)
    m_d.zero_(); step()
    _, pres = ctx.timed(step, max(1, min(steps, 2)), prepare=lambda: m_d.zero_())
    api.set_profile(False)
    sp_ms, sp_cnt = sum(r.info.spmv_ms for r in pres), sum(r.info.spmv_timed for r in pres)
    info = op.info()
    close_system(ctx, S)
    par = parity_block(ctx, kind, solver)
    if ctx.rank != 0:
        return None
    peak, _src = hbm_peak()
    bpi = bytes_per_iteration(solver, n, nnz)
    sp_alg = spmv_bytes(info["n_rows"], info["nnz"])
    sp_avg = sp_ms / max(sp_cnt, 1)
    return {"workload": name, "stencil": kind, "grid": g, "rows": n, "nnz": nnz, "solver": solver, "n_gpus": ctx.world,
            "value": value, "unit": "iterations/s", "steps": steps, "iterations_per_step": iters, "ms_per_step": ms / steps,
            "gpu_launches": sum(r.info.kernel_launches for r in res),
            "hbm_gbs_per_iteration_per_gpu": bpi * value / 1e9 / ctx.world, "iteration_frac_of_peak": bpi * value / 1e9 / ctx.world / peak,
            "iteration_frac_of_8TBps_nominal": bpi * value / 1e9 / ctx.world / 8000.0,
            "roofline": {"bound": "hbm", "kernel": "k_spmv (CSR SpMV + fused dots)", "achieved": sp_alg / (sp_avg * 1e-3) / 1e9 if sp_cnt else None,
                         "peak": peak, "unit": "GB/s", "frac": (sp_alg / (sp_avg * 1e-3) / 1e9 / peak) if sp_cnt else None, "traffic": None,
                         "algorithmic_bytes_per_launch": sp_alg, "avg_launch_ms": sp_avg, "launches_timed": sp_cnt},
            "parity": par}


def run_ours(args, wl):
    from liblcg_b200 import api, _lib

    kind, g, solver = wl
    sid = SOLVER_ID[solver]
    ctx = Ctx()
    torch, dev, world, rank, dist, stream = ctx.torch, ctx.dev, ctx.world, ctx.rank, ctx.dist, ctx.stream
    _lib.load()
    if args.poll > 0:
        api.set_poll_interval(args.poll)

    fixture = kind.startswith("fixture:")
    iters = args.iters
    para = api.lcg_default_parameters(epsilon=1e-300, max_iterations=iters)

    # ---- the system, generated on the device (rows of this rank)
    if fixture:
        if world > 1:
            raise SystemExit("the reference's 10K sample systems are single-GPU workloads")
        S0 = host_system(kind, g)
        n, nnz = S0["n"], S0["nnz"]
        S = dict(op=api.CsrOperator(S0["row_ptr"], S0["col"], S0["val"], jacobi=(solver == "PCG")), b=torch.from_numpy(S0["b"]).to(dev),
                 n=n, nnz=nnz, n_loc=n, part=None, transport=None, dev_csr=None)
    else:
        S = device_system(ctx, kind, g, solver, compress=args.compress, keep_csr=(world == 1))
        n, nnz = S["n"], S["nnz"]
    op, b_d, n_loc, transport, dev_csr = S["op"], S["b"], S["n_loc"], S["transport"], S["dev_csr"]
    m_d = torch.zeros(n_loc, dtype=torch.float64, device=dev)
    timed = ctx.timed

    def step_device():
        r = api.solve(op, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=stream)
        if r.ret != api.LCG_REACHED_MAX_ITERATIONS or r.iterations != iters:
            raise RuntimeError(f"solve returned {r.ret} after {r.iterations} iterations: {api.last_error()}")
        return r

    for _ in range(max(args.warmup, 3)):
        m_d.zero_()
        step_device()

    # ---- value: device-resident solve
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, res = timed(step_device, args.steps, prepare=lambda: m_d.zero_())
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(r.info.kernel_launches for r in res)
    dev_ms_inside = sum(r.info.device_ms for r in res)
    value = args.steps * iters / (ms_dev * 1e-3)

    # ---- per-iteration host synchronisation cost: the same solve with a trivial progress callback: the host
    # then makes one round trip per loop head, exactly like the reference's loop
    pf_value = None
    if world == 1:
        m_d.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rp_ = api.solve(op, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=stream, Pfp=lambda *a: 0)
        torch.cuda.synchronize()
        pf_value = {"value": rp_.iterations / (time.perf_counter() - t0), "unit": "iterations/s",
                    "note": "same step with a trivial progress callback: one host round trip per iteration"}

    # ---- roofline pass: same steps, every launch bracketed by events
    api.set_profile(True)
    m_d.zero_()
    step_device()
    _, pres = timed(step_device, args.steps, prepare=lambda: m_d.zero_())
    api.set_profile(False)
    spmv_ms = sum(r.info.spmv_ms for r in pres)
    spmv_cnt = sum(r.info.spmv_timed for r in pres)
    vec_ms = sum(r.info.vec_ms for r in pres)
    vec_cnt = sum(r.info.vec_timed for r in pres)
    prof_dev_ms = sum(r.info.device_ms for r in pres)

    # ---- e2e: HOST vectors in, HOST solution out, copies inside the timed region.  The headline uses page-locked host
    # buffers (what a caller who cares about PCIe allocates); the same steps with pageable numpy arrays are reported beside it.
    def e2e_leg(pinned):
        if pinned:
            m_h_t = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
            b_h_t = b_d.cpu().pin_memory()
            m_h, b_h = m_h_t.numpy(), b_h_t.numpy()
        else:
            m_h_t = b_h_t = None
            m_h, b_h = np.zeros(n_loc), b_d.cpu().numpy().copy()

        def zero_host():
            m_h[:] = 0.0   # the initial guess of the step, prepared outside the timed bracket

        if world == 1:
            def step_host():   # reference-shaped entry point with the sentinel callbacks
                if solver == "PCG":
                    rc = api.lcg_solver_preconditioned_cuda(api.CSR_AX, api.JACOBI_MX, None, m_h, b_h, n, nnz, para, op)
                else:
                    rc = api.lcg_solver_cuda(api.CSR_AX, None, m_h, b_h, n, nnz, para, op, solver_id=sid)
                if rc != api.LCG_REACHED_MAX_ITERATIONS:
                    raise RuntimeError(f"reference-shaped solve returned {rc}: {api.last_error()}")
                return float(m_h[0])   # the step's result is read on the host
            name = "lcg_solver_preconditioned_cuda" if solver == "PCG" else "lcg_solver_cuda"
        else:
            def step_host():   # partitioned solve: each rank hands its HOST slice of m, B to lcgb200_solve and reads m back
                r = api.solve(op, sid, m_h, b_h, param=para, device=False, jacobi=(solver == "PCG"), stream=stream)
                if r.ret != api.LCG_REACHED_MAX_ITERATIONS:
                    raise RuntimeError(f"partitioned host solve returned {r.ret}: {api.last_error()}")
                return float(m_h[0])
            name = "lcgb200_solve (host slices, row-partitioned handle)"
        zero_host()
        step_host()
        torch.cuda.synchronize()
        steps = args.steps if pinned else max(1, min(args.steps, 3))
        if world == 1:
            with torch.cuda.stream(torch.cuda.default_stream()):   # the reference-shaped calls run on the legacy default stream
                ms, _ = timed(step_host, steps, prepare=zero_host)
        else:
            ms, _ = timed(step_host, steps, prepare=zero_host)
        return steps * iters / (ms * 1e-3), ms / steps, name

    v_pin, ms_pin, api_name = e2e_leg(True)
    v_page, _, _ = e2e_leg(False)
    e2e = {"value": v_pin, "unit": "iterations/s", "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_pin,
           "api": api_name, "host_buffers": "page-locked (torch pin_memory)", "pageable_host_buffers": {"value": v_page, "unit": "iterations/s"}}

    info = op.info()
    fmt = op.format()

    # ---- the reference's OWN CUDA path on this GPU (cuBLAS host loop + cusparseSpMV callback, oracle/_ref/liblcg_ref_cuda.so)
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_ref_cuda and dev_csr is not None and solver in ("CG", "PCG", "CGS"):
        try:
            from oracle import pyoracle as po
            if po.have_reference_cuda():
                rc_lib = po.RefCuda()
                it_rc = min(iters, 100)
                mh = np.zeros(n)
                bh = b_d.cpu().numpy()
                rc_lib.solve(solver, n, nnz, dev_csr[0].data_ptr(), dev_csr[1].data_ptr(), dev_csr[2].data_ptr(), mh, bh, 1e-300, 5)   # warm-up (library init)
                best = None
                for _ in range(3):
                    mh[:] = 0.0
                    ret_rc, secs, _k = rc_lib.solve(solver, n, nnz, dev_csr[0].data_ptr(), dev_csr[1].data_ptr(), dev_csr[2].data_ptr(), mh, bh, 1e-300, it_rc)
                    best = secs if best is None else min(best, secs)
                # same system, same iteration count through our entry point: the two GPU paths must land on the same iterate
                mo = np.zeros(n)
                para_rc = api.lcg_default_parameters(epsilon=1e-300, max_iterations=it_rc)
                if solver == "PCG":
                    api.lcg_solver_preconditioned_cuda(api.CSR_AX, api.JACOBI_MX, None, mo, bh, n, nnz, para_rc, op)
                else:
                    api.lcg_solver_cuda(api.CSR_AX, None, mo, bh, n, nnz, para_rc, op, solver_id=sid)
                ref_cuda = {"value": it_rc / best, "unit": "iterations/s", "kind": "reference lcg_cuda.cu, unmodified: cuBLAS level-1 host loop + cusparseSpMV Ax callback"
                            + (" + lcg_vecDvecD_element_wise Jacobi Mx callback" if solver == "PCG" else ""),
                            "ret": ret_rc, "iterations": it_rc, "best_of": 3, "includes": "H2D of pageable m, B and D2H of m, like e2e.pageable_host_buffers",
                            "rel_l2_vs_ours_same_iterations": float(np.linalg.norm(mo - mh) / max(np.linalg.norm(mh), 1e-300)),
                            "speedup_e2e": v_page / (it_rc / best)}
        except Exception as exc:   # the baseline is optional evidence; never let it take the bench line down
            ref_cuda = {"unavailable": repr(exc)[:200]}
    # ---- the same steps on the optional compressed operator copy (LCGB200_CSR_COMPRESS): not the headline — `value` above is
    # plain CSR as SURVEY 8(d) defines the bytes — but what a caller gets by setting one flag on a matrix with few distinct rows
    compressed = None
    if rank == 0 and world == 1 and dev_csr is not None and not args.compress and not args.no_compressed_leg:
        try:
            opc = api.CsrOperator(dev_csr[0], dev_csr[1], dev_csr[2], jacobi=(solver == "PCG"), compress=True)
            cfmt = opc.format()
            if cfmt["level"] > 0:
                def step_c():
                    r = api.solve(opc, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=stream)
                    if r.ret != api.LCG_REACHED_MAX_ITERATIONS or r.iterations != iters:
                        raise RuntimeError(f"compressed solve returned {r.ret} after {r.iterations} iterations")
                    return r
                m_d.zero_(); step_c()
                ms_c, _ = timed(step_c, args.steps, prepare=lambda: m_d.zero_())
                x_c = m_d.clone()
                m_d.zero_(); step_device()
                diff = float(((x_c - m_d).norm() / m_d.norm()).item())
                api.set_profile(True)
                m_d.zero_(); pc = step_c()
                api.set_profile(False)
                sp_ms = pc.info.spmv_ms / max(pc.info.spmv_timed, 1)
                compressed = {"value": args.steps * iters / (ms_c * 1e-3), "unit": "iterations/s", "level": cfmt["level"],
                              "format": "row patterns: 1 byte per row" if cfmt["level"] == 2 else "dictionary codes: 2 bytes per entry",
                              "pattern_kernel": opc.pattern_kernel(),
                              "n_values": cfmt["n_values"], "n_offsets": cfmt["n_offsets"], "spmv_stream_bytes_per_launch": cfmt["stream_bytes"],
                              "spmv_avg_launch_ms": sp_ms, "spmv_stream_GBps": cfmt["stream_bytes"] / (sp_ms * 1e-3) / 1e9,
                              "spmv_frac_of_peak_on_format_bytes": cfmt["stream_bytes"] / (sp_ms * 1e-3) / 1e9 / hbm_peak()[0],
                              "rel_l2_vs_plain_csr_same_iterations": diff,
                              "note": "same solve, same entries; the SpMV streams the compressed copy instead of 12 bytes per non-zero"}
            opc.close()
        except Exception as exc:
            compressed = {"unavailable": repr(exc)[:200]}
    dev_csr = None
    S["dev_csr"] = None
    del m_d
    close_system(ctx, S)

    # ---- parity of this run's path (all ranks take part) and the other BASELINE configs, measured in the same run
    parity = None if fixture else parity_block(ctx, kind, solver, compress=args.compress)
    extras = {}
    if not fixture and not args.no_extra_legs:
        legs = []
        if args.workload == "pcg27_256":
            legs = (["cg7_128"] if world == 1 else []) + ["bicgstab7cd_512"]
        for name in legs:
            try:
                extras[name] = extra_leg(ctx, name, max(2, min(args.steps, 5)), 200 if name == "cg7_128" else 60)
            except Exception as exc:
                if world > 1:
                    raise   # a rank that drops out of a collective leg would hang the others
                extras[name] = {"unavailable": repr(exc)[:300]}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (SpMV + fused dot), per-rank rows
    peak, peak_src = hbm_peak()
    spmv_alg = spmv_bytes(info["n_rows"], info["nnz"])
    spmv_avg_ms = spmv_ms / max(spmv_cnt, 1)
    achieved = spmv_alg / (spmv_avg_ms * 1e-3) / 1e9 if spmv_cnt else None
    # dram bytes of ONE launch from an `ncu --set full` capture: only known for the configuration that was captured
    traffic, traffic_src = None, "not captured for this workload / GPU count"
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if world == 1 and not fmt["compressed"] and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if args.workload in tj:
                traffic, traffic_src = tj[args.workload], tj.get("source", "profiles/spmv_traffic.json")
        except Exception:
            pass
    bpi = bytes_per_iteration(solver, n, nnz)
    roofline = {"bound": "hbm", "kernel": "k_spmv<double, LPR, EpiDotAlpha> (CSR SpMV fused with the p.Ap dot)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_alg,
                "avg_launch_ms": spmv_avg_ms, "launches_timed": spmv_cnt,
                "operator_format": dict(fmt, note=("dictionary-compressed copy streamed (2 B per entry): `achieved` still counts the 12 B per entry of plain CSR, "
                                                   "so frac > 1 is the saved traffic, not missing work" if fmt["compressed"] else "plain CSR (12 B per entry)"),
                                        achieved_stream_GBps=(fmt["stream_bytes"] / (spmv_avg_ms * 1e-3) / 1e9 if spmv_cnt else None)),
                "share_of_step": spmv_ms / prof_dev_ms if prof_dev_ms else None,
                "vec_kernels": {"avg_launch_ms": vec_ms / max(vec_cnt, 1), "launches_timed": vec_cnt,
                                "share_of_step": vec_ms / prof_dev_ms if prof_dev_ms else None},
                "iteration": {"algorithmic_bytes": bpi, "achieved_GBps_per_gpu": bpi * value / 1e9 / world,
                              "frac_of_peak": bpi * value / 1e9 / world / peak, "frac_of_8TBps_nominal": bpi * value / 1e9 / world / 8000.0}}

    # ---- cpu baseline (N = 1 only): the reference CPU/OpenMP solver on the same system, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        t0 = time.time()
        rates, which, cores = cpu_reference_run(kind, g, solver, args.cpu_iters, repeats=1, warmup=0)
        cpu = {"value": rates[0], "unit": "iterations/s", "cores": cores, "kind": which,
               "sample": f"first {args.cpu_iters} {solver} iterations of the same {kind} {g}^3 system (max_iterations={args.cpu_iters}), "
                         f"OpenMP CSR Ax callback, {cores} threads; {time.time() - t0:.1f} s incl. host matrix generation"}

    line = {
        "metric": "cg_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "stencil": kind, "grid": g, "rows": n, "nnz": nnz, "solver": solver,
                   "iterations_per_step": iters, "parallelism": f"row-partition x{world}" if world > 1 else "single GPU", "transport": transport,
                   "l2": (f"inputs larger than L2: CSR {12 * nnz / world / 1e9:.2f} GB per GPU streamed every iteration (no flush needed)" if 12 * nnz / world > 126e6
                          else "cache-resident system: launch-latency-bound, it/s only (no roofline claim)"),
                   "lanes_per_row": info["lanes_per_row"], "tiles": info["n_tiles"], "operator_format": "dict-compressed" if fmt["compressed"] else "csr"},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "c3_cg7_128": extras.get("cg7_128"), "c5_bicgstab7cd_512": extras.get("bicgstab7cd_512"),
        "reference_cuda": ref_cuda, "compressed_operator": compressed, "with_progress_callback": pf_value, "clocks": clocks,
        "hbm_gbs_per_iteration": bpi * value / 1e9 / world,
        "diagnostics": {"solve_device_ms_per_step": dev_ms_inside / args.steps, "profile_pass_device_ms_per_step": prof_dev_ms / args.steps,
                        "kernel_ms_sum_per_step": (spmv_ms + vec_ms) / args.steps},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pcg27_256", choices=sorted(WORKLOADS))
    ap.add_argument("--iters", type=int, default=0, help="iterations per step (GPU arm); default 200 (60 for the 10K sample systems)")
    ap.add_argument("--cpu-iters", type=int, default=10, help="iterations of the cpu_baseline sample")
    ap.add_argument("--ref-iters", type=int, default=10, help="iterations per step of the --impl reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA (cuBLAS + cuSPARSE) leg")
    ap.add_argument("--compress", action="store_true", help="dictionary-compressed operator copy (LCGB200_CSR_COMPRESS): 2 bytes per entry streamed")
    ap.add_argument("--no-compressed-leg", action="store_true", help="skip the extra leg on the compressed operator copy")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the c3 (cg7_128) / c5 (bicgstab7cd_512) legs printed inside the default line")
    ap.add_argument("--poll", type=int, default=0, help="iterations enqueued per host poll of the convergence flag (0 = library default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.iters <= 0:
        args.iters = 60 if wl[0].startswith("fixture:") else 200
    if wl[0].startswith("fixture:"):
        args.cpu_iters = args.ref_iters = args.iters
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
