"""CPU tests (no GPU): the oracle restatement against the reference's golden vectors and, when the prebuilt
reference library is present, against the reference itself bit-for-bit.  Also pins the fixtures' known answers
(the reference's only "tests": sample8.cu:133-145,257, sample4.cpp:145-157, sample6.cpp:162-196)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from liblcg_b200 import io as lio, stencil

REAL = ["CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"]
CPLX = ["BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR"]
SETTINGS = {"eps1e-6": dict(epsilon=1e-6), "eps1e-10": dict(epsilon=1e-10), "eps1e-6_abs": dict(epsilon=1e-6, abs_diff=1)}
# iteration counts measured from the reference in SURVEY.md §8(c)
SURVEY_COUNTS = {"eps1e-6": [59, 57, 27, 36, 33, 71, 83], "eps1e-10": [100, 99, 57, 70, 56, 136, 217],
                 "eps1e-6_abs": [102, 100, 58, 72, 118, 136, 222]}


def check_against_golden(r, g, stride):
    assert r.ret == g["ret"]
    assert r.iters == g["iters"]
    assert r.calls == g["calls"]
    assert r.residual == pytest.approx(g["residual"], rel=1e-12, abs=0)
    assert np.linalg.norm(r.x) == pytest.approx(g["xnorm"], rel=1e-13)
    xs = r.x[::stride]
    if np.iscomplexobj(xs):
        np.testing.assert_allclose(xs.real, g["xs_re"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(xs.imag, g["xs_im"], rtol=1e-12, atol=1e-300)
    else:
        np.testing.assert_allclose(xs, g["xs"], rtol=1e-12, atol=1e-300)
    if "history" in g:
        np.testing.assert_allclose(r.history, g["history"], rtol=1e-12, atol=0)


def test_param_struct_layout(port):
    # lcg_para is 64 bytes with offsets 0/8/16/24/32/40/48/56, clcg_para 24 bytes (util.h:95-148, 247-273)
    import ctypes as C
    assert C.sizeof(po.LcgPara) == 64 and C.sizeof(po.ClcgPara) == 24
    offs = [getattr(po.LcgPara, f).offset for f, _ in po.LcgPara._fields_]
    assert offs == [0, 8, 16, 24, 32, 40, 48, 56]


def test_fixture_answers_are_exact(fixtures, port):
    # case_*_B are exact solutions of their systems: ||A ans - b|| / ||b|| ~ 1e-15
    A = fixtures["10K"]
    r = port.spmv(A, A["answer"]) - A["b"]
    assert np.linalg.norm(r) / np.linalg.norm(A["b"]) < 1e-13
    for name in ("10Kc", "1Kc"):
        Ac = fixtures[name]
        r = port.cspmv(Ac, Ac["answer"]) - Ac["b"]
        assert np.linalg.norm(r) / np.linalg.norm(Ac["b"]) < 1e-13


@pytest.mark.parametrize("setting", list(SETTINGS))
@pytest.mark.parametrize("sid", range(7))
def test_real_port_vs_golden(port, golden, fixtures, setting, sid):
    A = fixtures["10K"]
    n = A["n"]
    r = port.solve(sid, A, A["b"], para=po.default_para(**SETTINGS[setting]), low=np.full(n, -1e3), hig=np.full(n, 1e3),
                   diag=A["diag"], hist_cap=4096)
    check_against_golden(r, golden["real"][f"10K/{setting}/{REAL[sid]}"], golden["stride"])
    assert r.iters == SURVEY_COUNTS[setting][sid]


@pytest.mark.parametrize("k", [1, 10, 50])
@pytest.mark.parametrize("sid", range(7))
def test_real_port_pinned_iterations(port, golden, fixtures, k, sid):
    A = fixtures["10K"]
    n = A["n"]
    r = port.solve(sid, A, A["b"], para=po.default_para(epsilon=1e-300, max_iterations=k), low=np.full(n, -1e3),
                   hig=np.full(n, 1e3), diag=A["diag"], hist_cap=64)
    check_against_golden(r, golden["real"][f"10K/maxit{k}/{REAL[sid]}"], golden["stride"])
    assert r.ret == -1019  # LCG_REACHED_MAX_ITERATIONS


@pytest.mark.parametrize("sid", [5, 6])
def test_real_port_active_box(port, golden, fixtures, sid):
    A = fixtures["10K"]
    n = A["n"]
    r = port.solve(sid, A, A["b"], para=po.default_para(epsilon=1e-8, max_iterations=30), low=np.full(n, -10.0),
                   hig=np.full(n, 10.0), diag=A["diag"], hist_cap=512)
    check_against_golden(r, golden["real"][f"10K/box10/{REAL[sid]}"], golden["stride"])
    assert np.all(r.x <= 10.0) and np.all(r.x >= -10.0) and np.sum(np.abs(r.x) == 10.0) > 100


@pytest.mark.parametrize("fx", ["10Kc", "1Kc"])
@pytest.mark.parametrize("mode", ["abs", "rel"])
@pytest.mark.parametrize("sid", range(5))
def test_complex_port_vs_golden(port, golden, fixtures, fx, mode, sid):
    g = golden["complex"][f"{fx}/{mode}/{CPLX[sid]}"]
    if g["iters"] > 2000 and fx == "10Kc":
        pytest.skip("long BICGSTAB run is covered by the 1Kc fixture and the reference comparison")
    Ac = fixtures[fx]
    port.set_time(golden["seed"])
    r = port.csolve(sid, Ac, Ac["b"], para=po.default_cpara(abs_diff=1 if mode == "abs" else 0), hist_cap=20000)
    check_against_golden(r, g, golden["stride"])


@pytest.mark.parametrize("k", [1, 10, 50])
@pytest.mark.parametrize("sid", range(4))
def test_complex_port_pinned_iterations(port, golden, fixtures, k, sid):
    Ac = fixtures["10Kc"]
    port.set_time(golden["seed"])
    r = port.csolve(sid, Ac, Ac["b"], para=po.default_cpara(epsilon=1e-300, max_iterations=k), hist_cap=64)
    check_against_golden(r, golden["complex"][f"10Kc/maxit{k}/{CPLX[sid]}"], golden["stride"])
    assert r.ret == -1019  # the complex solvers return LCG_REACHED_MAX_ITERATIONS (clcg.cpp:164)


def test_complex_pcg_identity_equals_bicg_sym(port, golden, fixtures):
    # the reference has no buildable CPU complex PCG; with M = I it must reproduce clbicg_symmetric exactly
    Ac = fixtures["10Kc"]
    one = np.ones(Ac["n"], dtype=np.complex128)
    r = port.csolve(po.CLCG_PCG, Ac, Ac["b"], diag=one, para=po.default_cpara(abs_diff=1), hist_cap=20000)
    check_against_golden(r, golden["complex"]["10Kc/abs/BICG_SYM"], golden["stride"])


def test_tfqmr_max_iterations_terminates(port, fixtures):
    # deliberate fix of clcg.cpp:800-804 (the reference spins forever): returns LCG_REACHED_MAX_ITERATIONS
    Ac = fixtures["1Kc"]
    port.set_time(12345)
    r = port.csolve(po.CLCG_TFQMR, Ac, Ac["b"], para=po.default_cpara(epsilon=1e-300, max_iterations=7))
    assert r.ret == -1019 and r.iters == 7


@pytest.mark.parametrize("key", ["7pt/24/CG", "7pt/24/PCG", "7pt/24/CGS", "7pt/24/BICGSTAB", "27pt/16/CG", "27pt/16/PCG",
                                 "7pt_cd/20/CGS", "7pt_cd/20/BICGSTAB", "7pt_cd/20/BICGSTAB2"])
def test_stencil_port_vs_golden(port, golden, key):
    kind, g, name = key.split("/")
    S = stencil.make_system(kind, int(g))
    assert S["nnz"] == stencil.stencil_nnz(kind, int(g))
    d = lio.csr_diagonal(S["row_ptr"], S["col"], S["val"])
    r = port.solve(REAL.index(name), S, S["b"], para=po.default_para(epsilon=1e-10), diag=d, hist_cap=4096)
    check_against_golden(r, golden["stencil"][key], golden["stride"])
    assert np.linalg.norm(r.x - S["x_star"]) / np.linalg.norm(S["x_star"]) < 1e-3


def test_validation_order_and_codes(port, fixtures):
    A = fixtures["10K"]
    n = A["n"]
    lo, hi = np.full(n, -1.0), np.full(n, 1.0)
    bad = [(dict(max_iterations=-1), -1022), (dict(epsilon=0.0), -1021), (dict(epsilon=1.0), -1021)]
    for kw, code in bad:
        for sid in (0, 1, 2, 3):
            assert port.solve(sid, A, A["b"], para=po.default_para(**kw), diag=A["diag"]).ret == code
    # BICGSTAB2 reports epsilon >= 1 as INVILAD_RESTART_EPSILON (lcg.cpp:821-822), PG as INVALID_LAMBDA (lcg.cpp:1064-1065)
    assert port.solve(4, A, A["b"], para=po.default_para(epsilon=1.0)).ret == -1020
    assert port.solve(5, A, A["b"], para=po.default_para(epsilon=1.0), low=lo, hig=hi).ret == -1015
    assert port.solve(6, A, A["b"], para=po.default_para(sigma=1.5), low=lo, hig=hi).ret == -1014
    assert port.solve(6, A, A["b"], para=po.default_para(beta=0.0), low=lo, hig=hi).ret == -1013
    assert port.solve(6, A, A["b"], para=po.default_para(maxi_m=0), low=lo, hig=hi).ret == -1012


def test_already_optimised_and_stop(port, fixtures):
    A = fixtures["10K"]
    r = port.solve(0, A, A["b"], x0=A["answer"], para=po.default_para())
    assert r.ret == 2 and r.iters == 0 and r.calls == 1
    r = port.solve(0, A, A["b"], para=po.default_para(), stop_at=5)
    assert r.ret == 1 and r.iters == 5


# ---------------------------------------------------------------- live comparison with the reference library
@pytest.mark.parametrize("sid", range(7))
def test_real_port_equals_reference_bitwise(port, reflib, fixtures, sid):
    A = fixtures["10K"]
    n = A["n"]
    rng = np.random.default_rng(7 + sid)
    x0 = rng.standard_normal(n)   # warm start: exercises the in-place / non-zero initial guess path
    kw = dict(para=po.default_para(epsilon=1e-9, abs_diff=sid % 2, max_iterations=150), low=np.full(n, -50.0), hig=np.full(n, 50.0),
              diag=A["diag"], hist_cap=4096, x0=x0)
    a, b = port.solve(sid, A, A["b"], **kw), reflib.solve(sid, A, A["b"], **kw)
    assert (a.ret, a.iters, a.calls) == (b.ret, b.iters, b.calls)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.history, b.history)


@pytest.mark.parametrize("sid", range(5))
def test_complex_port_equals_reference_bitwise(port, reflib, fixtures, sid):
    Ac = fixtures["1Kc"]
    port.set_time(777)
    reflib.set_time(777)
    assert np.array_equal(port.vecrnd(64), reflib.vecrnd(64))
    kw = dict(para=po.default_cpara(epsilon=1e-6, abs_diff=1), hist_cap=30000)
    a, b = port.csolve(sid, Ac, Ac["b"], **kw), reflib.csolve(sid, Ac, Ac["b"], **kw)
    assert (a.ret, a.iters, a.calls) == (b.ret, b.iters, b.calls)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.history, b.history)


# ------------------------------------------------------------------------------------------------ IC(0)
def _lower_csr(A):
    rows = np.repeat(np.arange(A["n"]), np.diff(A["row_ptr"]))
    keep = A["col"] <= rows
    rp = np.zeros(A["n"] + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows[keep], minlength=A["n"]), out=rp[1:])
    return rp, A["col"][keep].astype(np.int32), A["val"][keep]


def test_ic0_factor_is_bit_identical_to_the_reference():
    """lcgb200_ic0_factor_host (liblcg_b200/csrc/ic0_host.h; host code, no GPU) against the reference's own factorisations compiled
    into oracle/_ref: lcg_incomplete_Cholesky_half_coo (preconditioner.cpp:33-160) on data/case_10K_A and
    clcg_incomplete_Cholesky_cuda_half (preconditioner_cuda.cu:40-270) in double and single complex on data/case_10K_cA — the same
    entries, bit for bit (the restatement performs the reference's operations in the reference's order)."""
    from liblcg_b200 import api, io as lio
    if not po.have_reference():
        pytest.skip("oracle/_ref not present")
    A = lio.load_fixture("10K")
    rp, ci, v = _lower_csr(A)
    ours = api.ic0_factor_host(rp, ci, v)
    ir, ic, iv = po.ref_ic0_half(A)
    assert np.array_equal(ic, ci) and np.array_equal(ir, np.repeat(np.arange(A["n"]), np.diff(rp)))
    assert np.array_equal(ours.view(np.int64), iv.view(np.int64))
    assert np.all(np.isfinite(ours))
    if po.have_reference_cuda():
        Ac = lio.load_fixture("10Kc")
        rpc, cic, vc = _lower_csr(Ac)
        for single in (False, True):
            dt = np.complex64 if single else np.complex128
            ours_c = api.ic0_factor_host(rpc, cic, vc.astype(dt))
            _, icc, ivc = po.ref_cic0_half(Ac, single=single)
            assert np.array_equal(icc, cic)
            assert ours_c.tobytes() == ivc.tobytes()
