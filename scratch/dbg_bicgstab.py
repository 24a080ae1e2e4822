import sys, numpy as np
sys.path.insert(0,'.')
from liblcg_b200 import api, io as lio
Ac = lio.load_fixture("1Kc")
op = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=True, jacobi=True)
for s in range(8):
    rng = np.random.default_rng(s)
    bp = Ac["b"] * (1 + (2e-16 if s else 0)*rng.standard_normal(Ac["n"]))
    api.set_shadow_seed(12345)
    hist = []
    m = np.zeros(Ac["n"], dtype=np.complex128)
    r = api.csolve(op, api.CLCG_BICGSTAB, m, bp, param=api.clcg_default_parameters(abs_diff=1, max_iterations=40000),
                   Pfp=lambda i, md, c, p, n, nz, k: hist.append(c) or 0)
    h = np.array(hist)
    print(s, "pf-mode ret", r.ret, "it", r.iterations, "max res %.3e" % np.nanmax(h), "last", h[-4:])
    m = np.zeros(Ac["n"], dtype=np.complex128)
    r = api.csolve(op, api.CLCG_BICGSTAB, m, bp, param=api.clcg_default_parameters(abs_diff=1, max_iterations=40000))
    print(s, "nopf    ret", r.ret, "it", r.iterations, "res", r.residual, "nan" if np.isnan(m).any() else "")
