"""Reference-order mode against the CPU port, case by case (run on a GPU box): return code, iteration count, the first
entry of the residual history that differs in its bits, and whether the solution is bit-identical.  Saves the GPU
histories and solutions to gpurun_out/exact_probe.npz for offline analysis."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po
from liblcg_b200 import api, io as lio

REAL = ["CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"]
CPLX = ["BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR", "PCG"]
SEED = 20240607
port = po.Oracle("port")
out = {}


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(a[:n].view(np.int64) != b[:n].view(np.int64))[0]
    return int(d[0]) if len(d) else (-1 if len(a) == len(b) else n)


def report(name, r, x, hist, cpu):
    hist = np.array(hist, dtype=np.float64)
    fd = first_diff(hist, cpu.history)
    same_x = np.array_equal(x, cpu.x)
    relx = float(np.linalg.norm(x - cpu.x) / max(np.linalg.norm(cpu.x), 1e-300))
    ok = r.ret == cpu.ret and r.iterations == cpu.iters and fd == -1 and same_x
    extra = ""
    if fd >= 0 and fd < min(len(hist), len(cpu.history)):
        extra = f" hist[{fd}] gpu {hist[fd].hex()} cpu {cpu.history[fd].hex()}"
    print(f"{'OK  ' if ok else 'DIFF'} {name:28s} ret {r.ret}/{cpu.ret} iters {r.iterations}/{cpu.iters} calls {len(hist)}/{cpu.calls} first_diff {fd} x_bitwise {same_x} relx {relx:.2e} ms {r.info.total_ms:.0f}{extra}", flush=True)
    out[name + "/hist"] = hist
    out[name + "/x"] = x
    return ok


def main():
    api.set_reference_order(True)
    fx = {k: lio.load_fixture(k) for k in ("10K", "1Kc", "10Kc")}
    good = total = 0
    A = fx["10K"]
    n = A["n"]
    diag = lio.csr_diagonal(A["row_ptr"], A["col"], A["val"])
    low, hig = np.full(n, -1e3), np.full(n, 1e3)
    settings = {"eps1e-6": dict(epsilon=1e-6), "eps1e-10": dict(epsilon=1e-10), "eps1e-6_abs": dict(epsilon=1e-6, abs_diff=1), "maxit10": dict(epsilon=1e-300, max_iterations=10)}
    only = sys.argv[1:] 
    for sname, kw in settings.items():
        for sid in range(7):
            name = f"10K/{sname}/{REAL[sid]}"
            if only and not any(o in name for o in only):
                continue
            hist = []
            op = api.CsrOperator(A["row_ptr"], A["col"], A["val"], jacobi=True)
            m = np.zeros(n)
            r = api.solve(op, sid, m, A["b"], low=low, hig=hig, param=api.lcg_default_parameters(**kw), jacobi=(sid == 1),
                          Pfp=lambda i, md, c, p, nn, nz, k: hist.append(c) or 0)
            op.close()
            cpu = port.solve(sid, A, A["b"], para=po.default_para(**kw), low=low, hig=hig, diag=diag, hist_cap=1 << 16)
            good += report(name, r, m, hist, cpu); total += 1
    csettings = {"abs": dict(abs_diff=1), "rel": dict(abs_diff=0), "maxit10": dict(epsilon=1e-300, max_iterations=10)}
    for f in ("10Kc", "1Kc"):
        Ac = fx[f]
        nc = Ac["n"]
        cdiag = lio.csr_diagonal(Ac["row_ptr"], Ac["col"], Ac["val"])
        for sname, kw in csettings.items():
            for sid in range(6):
                name = f"{f}/{sname}/{CPLX[sid]}"
                if only and not any(o in name for o in only):
                    continue
                pcg = CPLX[sid] == "PCG"
                api.set_shadow_seed(SEED); port.set_time(SEED)
                hist = []
                op = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=(sid == 0), jacobi=pcg)
                m = np.zeros(nc, dtype=np.complex128)
                r = api.csolve(op, api.CLCG_PCG if pcg else sid, m, Ac["b"], param=api.clcg_default_parameters(**kw), jacobi=pcg,
                               Pfp=lambda i, md, c, p, nn, nz, k: hist.append(c) or 0)
                op.close()
                cpu = port.csolve(po.CLCG_PCG if pcg else sid, Ac, Ac["b"], para=po.default_cpara(**kw), diag=cdiag if pcg else None, hist_cap=1 << 17)
                good += report(name, r, m, hist, cpu); total += 1
    print(f"{good}/{total} bit-identical")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "exact_probe.npz"), **out)


if __name__ == "__main__":
    main()
