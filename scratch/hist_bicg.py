import sys, numpy as np
sys.path.insert(0, '.')
from liblcg_b200 import api, io as lio
from oracle import pyoracle as po
Ac = lio.load_fixture("10Kc")
port = po.Oracle("port")
out = {}
for sid, nm in ((0, "BICG"), (1, "BICG_SYM")):
    hist = []
    op = api.CsrOperator(Ac["row_ptr"], Ac["col"], Ac["val"], transpose=True)
    m = np.zeros(Ac["n"], dtype=np.complex128)
    r = api.csolve(op, sid, m, Ac["b"], param=api.clcg_default_parameters(abs_diff=0), Pfp=lambda i, md, c, p, n, nz, k: hist.append(c) or 0)
    cpu = port.csolve(sid, Ac, Ac["b"], para=po.default_cpara(abs_diff=0), hist_cap=4000)
    g = np.array(hist); c = cpu.history
    print(nm, "gpu", r.iterations, "cpu", cpu.iters)
    for k in list(range(0, 60, 10)) + list(range(240, min(len(g), len(c)), 4)):
        print(k, "%.6e %.6e  ratio %.4f" % (g[k], c[k], g[k] / c[k]))
    # without callback too
    m = np.zeros(Ac["n"], dtype=np.complex128)
    r2 = api.csolve(op, sid, m, Ac["b"], param=api.clcg_default_parameters(abs_diff=0))
    print("no-callback iterations", r2.iterations)
