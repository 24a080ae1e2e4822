"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/lcgb200.h declares; parameter validation (which happens before any CUDA call) returns the reference's
integers in the reference's order; the product package never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lcgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(lcgb200_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from liblcg_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/lcgb200.h but not exported"
    # and the ctypes table covers the header (so tests call through typed prototypes)
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)


def test_struct_layouts_match_reference():
    from liblcg_b200._lib import LcgPara, ClcgPara
    assert C.sizeof(LcgPara) == 64 and C.sizeof(ClcgPara) == 24     # util.h:95-148, 247-273
    assert [getattr(LcgPara, f).offset for f, _ in LcgPara._fields_] == [0, 8, 16, 24, 32, 40, 48, 56]


def test_validation_happens_before_any_gpu_work():
    """Same order and integers as lcg_cuda.cu:91-98 / clcg_cuda.cu:94-101; none of these calls reaches CUDA."""
    from liblcg_b200 import api
    m = np.zeros(4)
    b = np.ones(4)
    P = api.lcg_default_parameters
    fake = 0x1000  # non-null "instance"/handles: validation must fail before they are dereferenced
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, b, 0, 0, P(), fake) == api.LCG_INVILAD_VARIABLE_SIZE
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, b, 4, 0, P(max_iterations=-1), fake) == api.LCG_INVILAD_MAX_ITERATIONS
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, b, 4, 0, P(epsilon=0.0), fake) == api.LCG_INVILAD_EPSILON
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, b, 4, 0, P(epsilon=1.0), fake) == api.LCG_INVILAD_EPSILON
    assert api.lcg_solver_cuda(api.CSR_AX, None, None, b, 4, 0, P(), fake) == api.LCG_INVALID_POINTER
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, None, 4, 0, P(), fake) == api.LCG_INVALID_POINTER
    # a user callback (not the sentinel) with null cuBLAS/cuSPARSE handles -> LCG_INVALID_POINTER (lcg_cuda.cu:97-98)
    user_cb = 0x2000
    assert api.lcg_solver_cuda(user_cb, None, m, b, 4, 0, P(), None, None, None) == api.LCG_INVALID_POINTER
    # constrained: lcg.cpp:1062-1070, 1232-1243
    lo, hi = -np.ones(4), np.ones(4)
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, lo, hi, 4, 0, P(epsilon=1.0), fake) == api.LCG_INVALID_LAMBDA
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, lo, hi, 4, 0, P(step=0.0), fake) == api.LCG_INVALID_LAMBDA
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, None, hi, 4, 0, P(), fake) == api.LCG_INVALID_POINTER
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, lo, hi, 4, 0, P(sigma=1.0), fake, solver_id=api.LCG_SPG) == api.LCG_INVALID_SIGMA
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, lo, hi, 4, 0, P(beta=1.0), fake, solver_id=api.LCG_SPG) == api.LCG_INVALID_BETA
    assert api.lcg_solver_constrained_cuda(api.CSR_AX, None, m, b, lo, hi, 4, 0, P(maxi_m=0), fake, solver_id=api.LCG_SPG) == api.LCG_INVALID_MAXIM
    # BICGSTAB2's oddly placed epsilon test (lcg.cpp:821-822)
    assert api.lcg_solver_cuda(api.CSR_AX, None, m, b, 4, 0, P(epsilon=1.0), fake, solver_id=api.LCG_BICGSTAB2) == api.LCG_INVILAD_RESTART_EPSILON
    # complex
    cm, cb = np.zeros(4, dtype=np.complex128), np.ones(4, dtype=np.complex128)
    CP = api.clcg_default_parameters
    assert api.clcg_solver_cuda(api.CSR_CAX, None, cm, cb, -1, 0, CP(), fake) == api.LCG_INVILAD_VARIABLE_SIZE
    assert api.clcg_solver_cuda(api.CSR_CAX, None, cm, cb, 4, 0, CP(epsilon=2.0), fake) == api.LCG_INVILAD_EPSILON
    assert api.clcg_solver_cuda(api.CSR_CAX, None, None, cb, 4, 0, CP(), fake) == api.CLCG_INVALID_POINTER
    assert api.clcg_solver_cuda(api.CSR_CAX, None, cm, cb, 4, 0, CP(), fake, solver_id=api.CLCG_PCG) == api.CLCG_UNKNOWN_SOLVER
    assert api.clcg_solver_preconditioned_cuda(api.CSR_CAX, api.JACOBI_CMX, None, cm, cb, 4, 0, CP(), fake, solver_id=api.CLCG_BICG) == api.CLCG_UNKNOWN_SOLVER


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under liblcg_b200/ may reference it (no CPU fallback)."""
    pkg = os.path.join(ROOT, "liblcg_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "lcg_oracle" not in text and "pyoracle" not in text and "liblcg_ref" not in text, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from liblcg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()


def test_no_gpu_means_error_not_fallback():
    """Without a CUDA device the compute entry points must report an error code, never a result."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from liblcg_b200 import api
    rp = np.array([0, 1, 2], dtype=np.int32)
    ci = np.array([0, 1], dtype=np.int32)
    v = np.array([2.0, 2.0])
    with pytest.raises(RuntimeError):
        api.CsrOperator(rp, ci, v)


def test_row_pattern_chains_and_tiling_on_cpu():
    """liblcg_b200/csrc/pat_host.h (stride choice, chains, subset masks of the row-pattern operator copy) and the work-item
    tiling of k_spmv_pat, emulated on the CPU by tests/cxx/pat_chain_check.cpp: every row written exactly once, no load out
    of range, y equal to the CSR product, for cubes / bricks / a row block with ghost columns / matrices that take the
    masked and the row-by-row paths."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cxx", "build", "pat_chain_check")
    r = subprocess.run(["make", "-C", os.path.join(root, "tests", "cxx"), exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:]


def test_dropin_coo_row2col_matches_the_reference_algorithm():
    """clcg_smZcoo_row2col (lcg_complex_cuda.cu:266-292; host code): the reference inserts every entry into a std::map keyed
    N * col + row and reads it back in key order with rows and columns exchanged — a repeated (row, col) keeps its last value and
    the output is shorter than nz.  The drop-in library's version (a stable sort) against that algorithm restated with a dict."""
    so = os.path.join(ROOT, "liblcg_b200", "liblcg_dropin.so")
    lib = C.CDLL(so)
    fn = getattr(lib, "_Z19clcg_smZcoo_row2colPKiS0_PK7double2iiPiS4_PS1_")
    rng = np.random.default_rng(5)
    N, nz = 37, 400
    keys = np.sort(rng.integers(0, N * N, size=nz))                # row-sorted COO with repeats
    row, col = (keys // N).astype(np.int32), (keys % N).astype(np.int32)
    val = (rng.standard_normal(nz) + 1j * rng.standard_normal(nz)).astype(np.complex128)
    table = {}
    for r, c, v in zip(row, col, val):
        table[int(N) * int(c) + int(r)] = v
    exp = sorted(table.items())
    orow, ocol, oval = np.full(nz, -1, np.int32), np.full(nz, -1, np.int32), np.zeros(nz, np.complex128)
    fn.restype = None
    fn(row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), C.c_int(N), C.c_int(nz),
       orow.ctypes.data_as(C.c_void_p), ocol.ctypes.data_as(C.c_void_p), oval.ctypes.data_as(C.c_void_p))
    k = len(exp)
    assert k < nz                                                  # the draw has repeats
    assert np.array_equal(orow[:k], [o // N for o, _ in exp]) and np.array_equal(ocol[:k], [o % N for o, _ in exp])
    assert np.array_equal(oval[:k], np.array([v for _, v in exp])) and np.all(orow[k:] == -1)


def test_cxx_dropin_headers_compile():
    """include/lcg_b200/{util,lcg_cuda,clcg_cuda}.h: a liblcg user's program (tests/cxx/dropin_sample.cu) compiles and links
    against them for sm_100a; a second translation unit checks the reference's names, values and default arguments."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "tests", "cxx")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(os.path.join(root, "tests", "cxx", "build", "dropin_sample"))
    src = r"""
#include "lcg_b200/lcg_cuda.h"
#include "lcg_b200/clcg_cuda.h"
#include "lcg_b200/lcg.h"
#include "lcg_b200/clcg.h"
#include "lcg_b200/solver_cuda.h"
#include "lcg_b200/algebra_cuda.h"
#include "lcg_b200/lcg_complex_cuda.h"
#include <cstddef>
static_assert(sizeof(lcg_para) == 64 && offsetof(lcg_para, epsilon) == 8 && offsetof(lcg_para, maxi_m) == 56, "lcg_para layout (util.h:95-148)");
static_assert(sizeof(clcg_para) == 24, "clcg_para layout (util.h:247-273)");
static_assert(LCG_SPG == 6 && CLCG_PBICG == 6 && LCG_REACHED_MAX_ITERATIONS == -1019 && LCG_SIZE_NOT_MATCH == -1011, "ids");
static_assert(CLCG_REACHED_MAX_ITERATIONS == -1020 && CLCG_NAN_VALUE == -1019 && CLCG_UNKNOWN_SOLVER == -1016, "complex ids");
int main() {
    lcg_para p = lcg_default_parameters(); clcg_para q = clcg_default_parameters();
    int (*f)(lcg_axfunc_cuda_ptr, lcg_progress_cuda_ptr, lcg_float*, const lcg_float*, const int, const int, const lcg_para*, void*,
             cublasHandle_t, cusparseHandle_t, lcg_solver_enum) = lcg_solver_cuda;
    int (*g)(clcg_axfunc_cuda_ptr, clcg_progress_cuda_ptr, cuDoubleComplex*, const cuDoubleComplex*, const int, const int, const clcg_para*,
             void*, cublasHandle_t, cusparseHandle_t, clcg_solver_enum) = clcg_solver_cuda;
    int (*hs)(lcg_axfunc_ptr, lcg_progress_ptr, lcg_float*, const lcg_float*, const int, const lcg_para*, void*, lcg_solver_enum) = lcg_solver;
    int (*hc)(clcg_axfunc_ptr, clcg_progress_ptr, lcg_complex*, const lcg_complex*, const int, const clcg_para*, void*, clcg_solver_enum) = clcg_solver;
    void (*dv)(const lcg_float*, const lcg_float*, lcg_float*, int, int) = lcg_vecDvecD_element_wise;       // algebra_cuda.h:84
    void (*gd)(const int*, const int*, const cuDoubleComplex*, const int, cuDoubleComplex*, int) = clcg_smZcsr_get_diagonal;   // lcg_complex_cuda.h:202
    if (!hs || !hc || !dv || !gd) return 2;
    return (p.epsilon == 1e-6 && q.epsilon == 1e-6 && f && g && lcg_select_solver("LCG_PG") == LCG_PG && lcg_select_solver("x") == LCG_CGS) ? 0 : 1;
}
"""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "names.cu")
        open(path, "w").write(src)
        exe = os.path.join(td, "names")
        r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-I", os.path.join(root, "include"), path, "-L", os.path.join(root, "liblcg_b200"),
                            "-llcgb200", "-Xlinker", "-rpath=" + os.path.join(root, "liblcg_b200"), "-o", exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        assert subprocess.run([exe]).returncode == 0


def test_read_case_matches_python_loader():
    """lcgb200_read_case (host only) against liblcg_b200/io.py on the reference's fixtures (data/README:1-10)."""
    import os
    import numpy as np
    from liblcg_b200 import api, io as lio
    for name, cx in (("case_10K_A", False), ("case_10K_cA", True), ("case_1K_cA", True)):
        path = os.path.join(lio.GOLDEN_DATA, name)
        c = api.read_case(path, cx)
        ref = lio.read_case(path, None, cx)
        assert c["n"] == ref["n"] and c["nnz"] == ref["nnz"]
        rp = np.zeros(c["n"] + 1, dtype=np.int64)
        np.add.at(rp, c["rows"].astype(np.int64) + 1, 1)
        assert np.array_equal(np.cumsum(rp), ref["row_ptr"]) and np.all(np.diff(c["rows"]) >= 0)
        assert np.array_equal(c["cols"], ref["col"]) and np.array_equal(c["vals"], ref["val"]) and np.array_equal(c["b"], ref["b"])
    import pytest
    with pytest.raises(RuntimeError):
        api.read_case(os.path.join(lio.GOLDEN_DATA, "does_not_exist"))


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints ONE JSON line with the driver's keys (tier contract) — on a tiny workload, CPU only."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "pcg27_64", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    # under torchrun only rank 0 runs it
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "pcg27_64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_dropin_library_exports_liblcg_cxx_symbols():
    """liblcg_dropin.so defines liblcg's OWN C++ symbols out of line (the mangled names a program built against the reference's
    headers references: entry points, lcg()/lcgs(), util / algebra helpers, the out-of-line members and vtables of the four
    wrapper classes), so an existing binary can be re-linked against it.  Checked through the dynamic symbol table (no GPU)."""
    import subprocess
    so = os.path.join(ROOT, "liblcg_b200", "liblcg_dropin.so")
    assert os.path.exists(so), "liblcg_dropin.so is missing: run __graft_entry__.build()"
    out = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    defined = {l.split(" ", 2)[2] for l in out.splitlines() if len(l.split(" ", 2)) == 3}
    para, cpara = "lcg_para const*", "clcg_para const*"
    ax = "void (*)(void*, double const*, double*, int)"
    pf = f"int (*)(void*, double const*, double, {para}, int, int)"
    cax = "void (*)(void*, cublasContext*, cusparseContext*, cusparseDnVecDescr*, cusparseDnVecDescr*, int, int)"
    cpf = f"int (*)(void*, double const*, double, {para}, int, int, int)"
    zax = "void (*)(void*, cublasContext*, cusparseContext*, cusparseDnVecDescr*, cusparseDnVecDescr*, int, int, cusparseOperation_t)"
    zpf = f"int (*)(void*, double2 const*, double, {cpara}, int, int, int)"
    hz = "std::complex<double>"
    must = [
        f"lcg_solver({ax}, {pf}, double*, double const*, int, {para}, void*, lcg_solver_enum)",
        f"lcg_solver_preconditioned({ax}, {ax}, {pf}, double*, double const*, int, {para}, void*, lcg_solver_enum)",
        f"lcg_solver_constrained({ax}, {pf}, double*, double const*, double const*, double const*, int, {para}, void*, lcg_solver_enum)",
        f"lcg({ax}, {pf}, double*, double const*, int, {para}, void*, double*, double*, double*)",
        f"lcgs({ax}, {pf}, double*, double const*, int, {para}, void*, double*, double*, double*, double*, double*, double*, double*)",
        f"clcg_solver(void (*)(void*, {hz} const*, {hz}*, int, lcg_matrix_e, clcg_complex_e), int (*)(void*, {hz} const*, double, {cpara}, int, int), "
        f"{hz}*, {hz} const*, int, {cpara}, void*, clcg_solver_enum)",
        f"lcg_solver_cuda({cax}, {cpf}, double*, double const*, int, int, {para}, void*, cublasContext*, cusparseContext*, lcg_solver_enum)",
        f"lcg_solver_preconditioned_cuda({cax}, {cax}, {cpf}, double*, double const*, int, int, {para}, void*, cublasContext*, cusparseContext*, lcg_solver_enum)",
        f"lcg_solver_constrained_cuda({cax}, {cpf}, double*, double const*, double const*, double const*, int, int, {para}, void*, cublasContext*, "
        "cusparseContext*, lcg_solver_enum)",
        f"clcg_solver_cuda({zax}, {zpf}, double2*, double2 const*, int, int, {cpara}, void*, cublasContext*, cusparseContext*, clcg_solver_enum)",
        f"clcg_solver_preconditioned_cuda({zax}, {zax}, {zpf}, double2*, double2 const*, int, int, {cpara}, void*, cublasContext*, cusparseContext*, clcg_solver_enum)",
        "lcg_default_parameters()", "clcg_default_parameters()", "lcg_error_str(int, bool)", "clcg_error_str(int, bool)",
        "lcg_malloc(int)", "lcg_free(double*)", "lcg_vecset(double*, double, int)", "lcg_dot(double&, double const*, double const*, int)",
        "clcg_malloc(int)", "clcg_inner(std::complex<double>&, std::complex<double> const*, std::complex<double> const*, int)",
        "LCG_Solver::LCG_Solver()", "LCG_Solver::Minimize(double*, double const*, int, lcg_solver_enum, bool, bool)",
        "CLCG_Solver::Minimize(std::complex<double>*, std::complex<double> const*, int, clcg_solver_enum, bool, bool)",
        "LCG_CUDA_Solver::MinimizePreconditioned(cublasContext*, cusparseContext*, double*, double*, int, int, lcg_solver_enum, bool, bool)",
        "CLCG_CUDA_Solver::Minimize(cublasContext*, cusparseContext*, double2*, double2*, int, int, clcg_solver_enum, bool, bool)",
        # algebra_cuda.h / lcg_complex_cuda.h: what the samples build their Jacobi Mx callbacks from (sample10.cu:117,193)
        "lcg_set2box_cuda(double const*, double const*, double*, int, bool, bool)",
        "lcg_smDcsr_get_diagonal(int const*, int const*, double const*, int, double*, int)",
        "lcg_vecMvecD_element_wise(double const*, double const*, double*, int, int)", "lcg_vecDvecD_element_wise(double const*, double const*, double*, int, int)",
        "clcg_smCcsr_get_diagonal(int const*, int const*, float2 const*, int, float2*, int)", "clcg_smZcsr_get_diagonal(int const*, int const*, double2 const*, int, double2*, int)",
        "clcg_vecMvecC_element_wise(float2 const*, float2 const*, float2*, int, int)", "clcg_vecMvecZ_element_wise(double2 const*, double2 const*, double2*, int, int)",
        "clcg_vecDvecC_element_wise(float2 const*, float2 const*, float2*, int, int)", "clcg_vecDvecZ_element_wise(double2 const*, double2 const*, double2*, int, int)",
        "clcg_vecC_conjugate(float2 const*, float2*, int, int)", "clcg_vecZ_conjugate(double2 const*, double2*, int, int)",
        "clcg_smCcoo_row2col(int const*, int const*, float2 const*, int, int, int*, int*, float2*)",
        "clcg_smZcoo_row2col(int const*, int const*, double2 const*, int, int, int*, int*, double2*)",
        "cuda2lcg_complex(double2)", "lcg2cuda_complex(std::complex<double>)", "clcg_malloc_cuda(unsigned long)", "clcg_free_cuda(double2*)",
        "clcg_vecset_cuda(double2*, double2, unsigned long)", "clcg_Cscale(float, float2)", "clcg_Csum(float2, float2)", "clcg_Cdiff(float2, float2)",
        "clcg_Csqrt(float2)", "clcg_Zscale(double, double2)", "clcg_Zsum(double2, double2)", "clcg_Zdiff(double2, double2)", "clcg_Zsqrt(double2)",
    ]
    missing = [m for m in must if m not in defined]
    assert not missing, missing


def test_reference_order_build_shares_no_kernel_with_the_fast_build():
    """The solver sources are compiled twice (fast build / reference-order build with -fmad=false).  A kernel instantiated
    under the same name in both would be an ODR collision: the linker keeps one host stub and one of the two device images
    would silently serve both builds.  The variant tag of the Engine launch helpers must keep the kernel sets disjoint, and the
    reference-order objects must contain only the exact.cuh kernels."""
    import shutil
    build = os.path.join(ROOT, "liblcg_b200", "csrc", "build")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.isdir(build) and os.path.exists(cuobjdump)):
        pytest.skip("object files / cuobjdump not available")

    def kernels(obj):
        out = subprocess.run([cuobjdump, "-elf", os.path.join(build, obj)], capture_output=True, text=True).stdout
        return set(re.findall(r"\.text\.(_Z\w+)", out))

    exact = kernels("solvers_real_x.o") | kernels("solvers_complex_x.o")
    fast = set()
    for o in ("solvers_real.o", "solvers_complex.o", "solvers_complexf.o", "capi.o", "comm.o", "engine.o", "datastep.o", "stencil_gen.o"):
        fast |= kernels(o)
    assert exact and fast
    assert not (exact & fast), sorted(exact & fast)[:5]
    assert all(re.match(r"_ZN7lcgb200(6kx_vec|7kx_spmv|8kx_total)I", k) for k in exact), [k for k in exact if "kx_" not in k][:5]
