"""TEST INFRASTRUCTURE — regenerate tests/golden/golden.json from the UNMODIFIED reference library.

Run in the build container (needs oracle/_ref/liblcg_ref.so, i.e. /root/reference at build time):
    python oracle/make_golden.py
The GPU box never runs this; tests read the committed JSON.

Recorded per case: return code, iteration count (k of the last progress call), number of progress calls, final
residual, ||x||_2, every 499th solution component, and (for short runs) the whole residual history.
Cases = the reference's own fixtures (data/case_10K_A, case_10K_cA, case_1K_cA; sample8.cu:133-145,241-243;
sample6.cpp:162-196; sample4.cpp:145-157) under the three parameter settings SURVEY.md §8(c) lists, plus small
synthetic stencils of §8(d).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from liblcg_b200 import io as lio, stencil  # noqa: E402

SEED = 12345
STRIDE = 499


def rec(r, with_hist=True):
    d = dict(ret=int(r.ret), iters=int(r.iters), calls=int(r.calls), residual=float(r.residual),
             xnorm=float(np.linalg.norm(r.x)))
    xs = r.x[::STRIDE]
    if np.iscomplexobj(xs):
        d["xs_re"] = [float(v) for v in xs.real]
        d["xs_im"] = [float(v) for v in xs.imag]
    else:
        d["xs"] = [float(v) for v in xs]
    if with_hist and len(r.history) <= 400:
        d["history"] = [float(v) for v in r.history]
    return d


def main():
    ref = po.Oracle("reference")
    out = {"seed": SEED, "stride": STRIDE, "real": {}, "complex": {}, "stencil": {}}
    real_names = ["CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"]
    A = lio.load_fixture("10K")
    diag = lio.csr_diagonal(A["row_ptr"], A["col"], A["val"])
    n = A["n"]
    low, hig = np.full(n, -1e3), np.full(n, 1e3)
    settings = {"eps1e-6": dict(epsilon=1e-6), "eps1e-10": dict(epsilon=1e-10), "eps1e-6_abs": dict(epsilon=1e-6, abs_diff=1)}
    for sname, kw in settings.items():
        for sid, nm in enumerate(real_names):
            r = ref.solve(sid, A, A["b"], para=po.default_para(**kw), low=low, hig=hig, diag=diag, hist_cap=4096)
            out["real"][f"10K/{sname}/{nm}"] = rec(r)
    # pinned iteration counts: both sides stop on max_iterations, compare x after exactly k steps
    for k in (1, 10, 50):
        for sid, nm in enumerate(real_names):
            r = ref.solve(sid, A, A["b"], para=po.default_para(epsilon=1e-300, max_iterations=k), low=low, hig=hig, diag=diag, hist_cap=64)
            out["real"][f"10K/maxit{k}/{nm}"] = rec(r)
    # tight box: the projection is active (exercises lcg_set2box)
    lo2, hi2 = np.full(n, -10.0), np.full(n, 10.0)
    for sid in (5, 6):
        r = ref.solve(sid, A, A["b"], para=po.default_para(epsilon=1e-8, max_iterations=30), low=lo2, hig=hi2, diag=diag, hist_cap=512)
        out["real"][f"10K/box10/{real_names[sid]}"] = rec(r)
    cnames = ["BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR"]
    for fx in ("10Kc", "1Kc"):
        Ac = lio.load_fixture(fx)
        for sname, kw in {"abs": dict(abs_diff=1), "rel": dict(abs_diff=0)}.items():
            for sid, nm in enumerate(cnames):
                ref.set_time(SEED)
                r = ref.csolve(sid, Ac, Ac["b"], para=po.default_cpara(**kw), hist_cap=20000)
                out["complex"][f"{fx}/{sname}/{nm}"] = rec(r)
        for k in (1, 10, 50):
            for sid, nm in enumerate(cnames):
                if nm == "TFQMR":
                    continue  # the reference never terminates on max_iterations there (clcg.cpp:800-804)
                ref.set_time(SEED)
                r = ref.csolve(sid, Ac, Ac["b"], para=po.default_cpara(epsilon=1e-300, max_iterations=k), hist_cap=64)
                out["complex"][f"{fx}/maxit{k}/{nm}"] = rec(r)
    # small stencils (SURVEY.md §8(d) generators)
    for kind, g, sids in (("7pt", 24, (0, 1, 2, 3)), ("27pt", 16, (0, 1)), ("7pt_cd", 20, (2, 3, 4))):
        S = stencil.make_system(kind, g)
        d = lio.csr_diagonal(S["row_ptr"], S["col"], S["val"])
        for sid in sids:
            r = ref.solve(sid, S, S["b"], para=po.default_para(epsilon=1e-10), diag=d, hist_cap=4096)
            out["stencil"][f"{kind}/{g}/{real_names[sid]}"] = rec(r)
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", sum(len(v) for v in out.values() if isinstance(v, dict)), "cases")


if __name__ == "__main__":
    main()
