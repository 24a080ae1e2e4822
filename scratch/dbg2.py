import sys, numpy as np
sys.path.insert(0,'.')
from liblcg_b200 import api, io as lio
Ac = lio.load_fixture("1Kc")
op = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=True, jacobi=True)
api.set_shadow_seed(12345)
m = np.zeros(Ac["n"], dtype=np.complex128)
r = api.csolve(op, api.CLCG_BICGSTAB, m, Ac["b"], param=api.clcg_default_parameters(abs_diff=1, max_iterations=40000),
               Pfp=lambda i, md, c, p, n, nz, k: 0)
print(r.ret, r.iterations)
