# small memcheck workload: every kernel family once (real + complex, ragged + long rows, stencils) on tiny systems
import sys, numpy as np
sys.path.insert(0, '.')
from liblcg_b200 import api, stencil, io as lio
rng = np.random.default_rng(1)
def ragged(n, cx=False):
    lens = rng.integers(0, 9, size=n); lens[5] = 2300
    rp = np.zeros(n + 1, dtype=np.int32); np.cumsum(lens, out=rp[1:])
    col = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in lens]).astype(np.int32)
    val = rng.standard_normal(len(col)) + (1j * rng.standard_normal(len(col)) if cx else 0)
    d = np.arange(n)
    return rp, col, val
for kind, g in (("7pt", 9), ("27pt", 7), ("7pt_cd", 8)):
    S = stencil.make_system(kind, g)
    op = api.CsrOperator(S["row_ptr"], S["col"], S["val"], jacobi=True)
    for sid in range(7):
        m = np.zeros(S["n"]); lo = np.full(S["n"], -5.0); hi = np.full(S["n"], 5.0)
        r = api.solve(op, sid, m, S["b"], low=lo, hig=hi, param=api.lcg_default_parameters(epsilon=1e-8, max_iterations=12), jacobi=(sid == 1))
    op.close()
rp, col, val = ragged(3000)
op = api.CsrOperator(rp, col, val)
import torch
x = torch.randn(3000, dtype=torch.float64, device="cuda"); y = torch.empty_like(x); d = torch.zeros(3, dtype=torch.float64, device="cuda")
op.spmv(x, y); op.spmv_dot(x, y, None, d); torch.cuda.synchronize(); op.close()
Ac = lio.load_fixture("1Kc")
op = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=True, jacobi=True)
api.set_shadow_seed(3)
for sid in range(6):
    m = np.zeros(Ac["n"], dtype=np.complex128)
    api.csolve(op, sid, m, Ac["b"], param=api.clcg_default_parameters(max_iterations=10), jacobi=(sid == 5))
op.close()
print("sanitize workload done")
