"""Parity report (run on the GPU box): for every solver and fixture prints GPU vs CPU-oracle iteration counts and
solution differences next to the CPU oracle's OWN sensitivity to a 1-ulp perturbation of b — the yardstick that
tells a kernel bug (difference >> sensitivity) from the rounding sensitivity of an erratic recurrence.

    python tests/parity_report.py > profiles/parity_rNN.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from liblcg_b200 import api, io as lio, stencil  # noqa: E402

REAL = ["CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"]
CPLX = ["BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR", "PCG"]


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    port = po.Oracle("port")
    rng = np.random.default_rng(0)
    A = lio.load_fixture("10K")
    n = A["n"]
    diag = lio.csr_diagonal(A["row_ptr"], A["col"], A["val"])
    low, hig = np.full(n, -1e3), np.full(n, 1e3)
    bp = A["b"] * (1 + 2e-16 * rng.standard_normal(n))
    op = api.CsrOperator(A["row_ptr"], A["col"], A["val"], jacobi=True)
    print("# real, case_10K_A: solver | setting | gpu_it cpu_it cpu_it(b+1ulp) | rel(x_gpu,x_cpu) rel(x_cpu',x_cpu)")
    settings = [("maxit1", dict(epsilon=1e-300, max_iterations=1)), ("maxit10", dict(epsilon=1e-300, max_iterations=10)),
                ("maxit50", dict(epsilon=1e-300, max_iterations=50)), ("eps1e-6", dict(epsilon=1e-6)),
                ("eps1e-10", dict(epsilon=1e-10)), ("eps1e-6_abs", dict(epsilon=1e-6, abs_diff=1))]
    for sid, nm in enumerate(REAL):
        for sname, kw in settings:
            m = np.zeros(n)
            r = api.solve(op, sid, m, A["b"], low=low, hig=hig, param=api.lcg_default_parameters(**kw), jacobi=(sid == 1))
            c = port.solve(sid, A, A["b"], para=po.default_para(**kw), low=low, hig=hig, diag=diag)
            c2 = port.solve(sid, A, bp, para=po.default_para(**kw), low=low, hig=hig, diag=diag)
            print(f"{nm:10s} {sname:12s} ret {r.ret:6d}/{c.ret:6d}  it {r.iterations:5d} {c.iters:5d} {c2.iters:5d}   "
                  f"dx {rel(m, c.x):.2e}  sens {rel(c2.x, c.x):.2e}")
    op.close()
    for fx in ("10Kc", "1Kc"):
        Ac = lio.load_fixture(fx)
        nc = Ac["n"]
        dg = lio.csr_diagonal(Ac["row_ptr"], Ac["col"], Ac["val"])
        bpc = Ac["b"] * (1 + 2e-16 * rng.standard_normal(nc))
        opc = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=True, jacobi=True)
        print(f"# complex, case_{fx}: solver | setting | gpu_it cpu_it cpu_it(b+1ulp) | rel(x_gpu,x_cpu) rel(x_cpu',x_cpu)")
        csettings = [("maxit1", dict(epsilon=1e-300, max_iterations=1)), ("maxit10", dict(epsilon=1e-300, max_iterations=10)),
                     ("maxit50", dict(epsilon=1e-300, max_iterations=50)), ("abs", dict(abs_diff=1)), ("rel", dict(abs_diff=0))]
        for sid, nm in enumerate(CPLX):
            for sname, kw in csettings:
                api.set_shadow_seed(12345)
                m = np.zeros(nc, dtype=np.complex128)
                r = api.csolve(opc, sid, m, Ac["b"], param=api.clcg_default_parameters(**kw), jacobi=(nm == "PCG"))
                port.set_time(12345)
                c = port.csolve(sid, Ac, Ac["b"], diag=dg, para=po.default_cpara(**kw))
                port.set_time(12345)
                c2 = port.csolve(sid, Ac, bpc, diag=dg, para=po.default_cpara(**kw))
                print(f"{nm:10s} {sname:12s} ret {r.ret:6d}/{c.ret:6d}  it {r.iterations:5d} {c.iters:5d} {c2.iters:5d}   "
                      f"dx {rel(m, c.x):.2e}  sens {rel(c2.x, c.x):.2e}")
        opc.close()
    print("# stencils (SURVEY §8(d)), eps 1e-10: kind g solver | gpu_it cpu_it | rel(x_gpu,x_cpu) rel(x_cpu',x_cpu)")
    for kind, g, sids in (("7pt", 48, (0, 1, 2, 3, 4)), ("27pt", 40, (0, 1, 2, 3)), ("7pt_cd", 48, (2, 3, 4))):
        S = stencil.make_system(kind, g)
        d = lio.csr_diagonal(S["row_ptr"], S["col"], S["val"])
        bps = S["b"] * (1 + 2e-16 * rng.standard_normal(S["n"]))
        ops = api.CsrOperator(S["row_ptr"], S["col"], S["val"], jacobi=True)
        for sid in sids:
            m = np.zeros(S["n"])
            r = api.solve(ops, sid, m, S["b"], param=api.lcg_default_parameters(epsilon=1e-10), jacobi=(sid == 1))
            c = port.solve(sid, S, S["b"], para=po.default_para(epsilon=1e-10), diag=d)
            c2 = port.solve(sid, S, bps, para=po.default_para(epsilon=1e-10), diag=d)
            print(f"{kind:7s} {g:3d} {REAL[sid]:10s} ret {r.ret}/{c.ret}  it {r.iterations:5d} {c.iters:5d} {c2.iters:5d}   "
                  f"dx {rel(m, c.x):.2e}  sens {rel(c2.x, c.x):.2e}")
        ops.close()


if __name__ == "__main__":
    main()
