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
def run_ours(args, wl):
    import torch
    from liblcg_b200 import api, _lib

    kind, g, solver = wl
    sid = SOLVER_ID[solver]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (liblcg_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    if args.poll > 0:
        api.set_poll_interval(args.poll)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    fixture = kind.startswith("fixture:")
    if fixture:
        if world > 1:
            raise SystemExit("the reference's 10K sample systems are single-GPU workloads")
        S0 = host_system(kind, g)
        n, nnz = S0["n"], S0["nnz"]
    else:
        n, nnz = g ** 3, stencil_nnz(kind, g)
    transport = None
    dev_csr = None
    iters = args.iters
    para = api.lcg_default_parameters(epsilon=1e-300, max_iterations=iters)

    # ---- the system, generated on the device (rows of this rank)
    if world > 1:
        from liblcg_b200 import dist as ldist
        part = ldist.build_stencil_partition(kind, g, rank, world, dev, jacobi=(solver == "PCG"), compress=args.compress)
        op, b_d, n_loc = part.op, part.b, part.n_local
        transport = "nvlink-p2p (halo + reduction totals pushed into peer memory from inside the kernels)" if part.p2p else "nccl (send/recv halo + allreduce)"
    elif fixture:
        op = api.CsrOperator(S0["row_ptr"], S0["col"], S0["val"], jacobi=(solver == "PCG"))
        b_d = torch.from_numpy(S0["b"]).to(dev)
        n_loc = n
    else:
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
        op = api.CsrOperator(rp, ci, va, jacobi=(solver == "PCG"), compress=args.compress)
        dev_csr = (rp, ci, va)   # kept for the reference-CUDA leg (the operator owns its own copy)
        n_loc = n
    m_d = torch.zeros(n_loc, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        r = api.solve(op, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=stream)
        if r.ret != api.LCG_REACHED_MAX_ITERATIONS or r.iterations != iters:
            raise RuntimeError(f"solve returned {r.ret} after {r.iterations} iterations: {api.last_error()}")
        return r

    def timed(fn, steps, prepare=None):
        """K steps, each bracketed by barrier + synchronize and a CUDA-event pair on the launching stream; the step times
        are summed and the MAX over ranks is taken.  `prepare` (resetting the initial guess) runs between the brackets."""
        total, res = 0.0, []
        for _ in range(steps):
            if prepare is not None:
                prepare()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res.append(fn())
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        barrier()
        return total, res

    for _ in range(max(args.warmup, 3)):
        m_d.zero_()
        step_device()

    # ---- value: device-resident solve
    sampler = ClockSampler(local_rank)
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

    # ---- e2e: reference-shaped entry point, pinned HOST vectors, copies inside the timed region
    e2e = None
    if world == 1:
        m_h_t = torch.zeros(n, dtype=torch.float64).pin_memory()
        b_h_t = b_d.cpu().pin_memory()
        m_h, b_h = m_h_t.numpy(), b_h_t.numpy()

        def zero_host():
            m_h[:] = 0.0   # the initial guess of the step, prepared outside the timed bracket

        def step_host():
            if solver == "PCG":
                rc = api.lcg_solver_preconditioned_cuda(api.CSR_AX, api.JACOBI_MX, None, m_h, b_h, n, nnz, para, op)
            else:
                rc = api.lcg_solver_cuda(api.CSR_AX, None, m_h, b_h, n, nnz, para, op, solver_id=sid)
            if rc != api.LCG_REACHED_MAX_ITERATIONS:
                raise RuntimeError(f"reference-shaped solve returned {rc}: {api.last_error()}")
            return float(m_h[0])   # the step's result is read on the host

        zero_host()
        step_host()
        torch.cuda.synchronize()
        with torch.cuda.stream(torch.cuda.default_stream()):   # the reference-shaped calls run on the legacy default stream
            ms_e2e, _ = timed(step_host, args.steps, prepare=zero_host)
        e2e = {"value": args.steps * iters / (ms_e2e * 1e-3), "unit": "iterations/s",
               "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_e2e / args.steps,
               "api": "lcg_solver_preconditioned_cuda" if solver == "PCG" else "lcg_solver_cuda"}
    else:
        # partitioned solve: each rank hands its HOST slice of m, B to lcgb200_solve (host vectors) and reads m back
        m_h_t = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
        b_h_t = b_d.cpu().pin_memory()
        m_h, b_h = m_h_t.numpy(), b_h_t.numpy()

        def zero_host():
            m_h[:] = 0.0

        def step_host():
            r = api.solve(op, sid, m_h, b_h, param=para, device=False, jacobi=(solver == "PCG"), stream=stream)
            if r.ret != api.LCG_REACHED_MAX_ITERATIONS:
                raise RuntimeError(f"partitioned host solve returned {r.ret}: {api.last_error()}")
            return float(m_h[0])

        zero_host()
        step_host()
        ms_e2e, _ = timed(step_host, args.steps, prepare=zero_host)
        e2e = {"value": args.steps * iters / (ms_e2e * 1e-3), "unit": "iterations/s",
               "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_e2e / args.steps,
               "api": "lcgb200_solve (host slices, row-partitioned handle)"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (SpMV + fused dot), per-rank rows
    peaks, peak_src = None, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        peak = 6650.0
        peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    info = op.info()
    spmv_alg = spmv_bytes(info["n_rows"], info["nnz"])
    spmv_avg_ms = spmv_ms / max(spmv_cnt, 1)
    achieved = spmv_alg / (spmv_avg_ms * 1e-3) / 1e9 if spmv_cnt else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    bpi = bytes_per_iteration(solver, n, nnz)
    roofline = {"bound": "hbm", "kernel": "k_spmv<double, LPR, EpiDotAlpha> (CSR SpMV fused with the p.Ap dot)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_alg,
                "avg_launch_ms": spmv_avg_ms, "launches_timed": spmv_cnt,
                "operator_format": dict(op.format(), note=("dictionary-compressed copy streamed (2 B per entry): `achieved` still counts the 12 B per entry of plain CSR, "
                                                             "so frac > 1 is the saved traffic, not missing work" if op.format()["compressed"] else "plain CSR (12 B per entry)"),
                                        achieved_stream_GBps=(op.format()["stream_bytes"] / (spmv_avg_ms * 1e-3) / 1e9 if spmv_cnt else None)),
                "share_of_step": spmv_ms / prof_dev_ms if prof_dev_ms else None,
                "vec_kernels": {"avg_launch_ms": vec_ms / max(vec_cnt, 1), "launches_timed": vec_cnt,
                                "share_of_step": vec_ms / prof_dev_ms if prof_dev_ms else None},
                "iteration": {"algorithmic_bytes": bpi, "achieved_GBps_per_gpu": bpi * value / 1e9 / world,
                              "frac_of_peak": bpi * value / 1e9 / world / peak, "frac_of_8TBps_nominal": bpi * value / 1e9 / world / 8000.0}}

    # ---- the reference's OWN CUDA path on this GPU (cuBLAS host loop + cusparseSpMV callback, oracle/_ref/liblcg_ref_cuda.so)
    ref_cuda = None
    if world == 1 and not args.no_ref_cuda and dev_csr is not None and solver in ("CG", "PCG", "CGS"):
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
                            "ret": ret_rc, "iterations": it_rc, "best_of": 3, "includes": "H2D of m, B and D2H of m, like e2e",
                            "rel_l2_vs_ours_same_iterations": float(np.linalg.norm(mo - mh) / max(np.linalg.norm(mh), 1e-300)),
                            "speedup_e2e": e2e["value"] / (it_rc / best)}
        except Exception as exc:   # the baseline is optional evidence; never let it take the bench line down
            ref_cuda = {"unavailable": repr(exc)[:200]}
    # ---- the same steps on the optional compressed operator copy (LCGB200_CSR_COMPRESS): not the headline — `value` above is
    # plain CSR as SURVEY 8(d) defines the bytes — but what a caller gets by setting one flag on a matrix with few distinct rows
    compressed = None
    if world == 1 and dev_csr is not None and not args.compress and not args.no_compressed_leg:
        try:
            opc = api.CsrOperator(dev_csr[0], dev_csr[1], dev_csr[2], jacobi=(solver == "PCG"), compress=True)
            fmt = opc.format()
            if fmt["level"] > 0:
                def step_c():
                    r = api.solve(opc, sid, m_d, b_d, param=para, device=True, jacobi=(solver == "PCG"), stream=stream)
                    if r.ret != api.LCG_REACHED_MAX_ITERATIONS or r.iterations != iters:
                        raise RuntimeError(f"compressed solve returned {r.ret} after {r.iterations} iterations")
                    return r
                m_d.zero_(); step_c()
                x_plain = None
                ms_c, _ = timed(step_c, args.steps, prepare=lambda: m_d.zero_())
                x_c = m_d.clone()
                m_d.zero_(); step_device()
                diff = float(((x_c - m_d).norm() / m_d.norm()).item())
                api.set_profile(True)
                m_d.zero_(); pc = step_c()
                api.set_profile(False)
                sp_ms = pc.info.spmv_ms / max(pc.info.spmv_timed, 1)
                compressed = {"value": args.steps * iters / (ms_c * 1e-3), "unit": "iterations/s", "level": fmt["level"],
                              "format": "row patterns: 1 byte per row" if fmt["level"] == 2 else "dictionary codes: 2 bytes per entry",
                              "n_values": fmt["n_values"], "n_offsets": fmt["n_offsets"], "spmv_stream_bytes_per_launch": fmt["stream_bytes"],
                              "spmv_avg_launch_ms": sp_ms, "spmv_stream_GBps": fmt["stream_bytes"] / (sp_ms * 1e-3) / 1e9,
                              "rel_l2_vs_plain_csr_same_iterations": diff,
                              "note": "same solve, same entries; the SpMV streams the compressed copy instead of 12 bytes per non-zero"}
            opc.close()
        except Exception as exc:
            compressed = {"unavailable": repr(exc)[:200]}
    dev_csr = None
    torch.cuda.empty_cache()

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
                   "lanes_per_row": info["lanes_per_row"], "tiles": info["n_tiles"], "operator_format": "dict-compressed" if op.format()["compressed"] else "csr"},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": ref_cuda, "compressed_operator": compressed, "with_progress_callback": pf_value, "clocks": clocks,
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
