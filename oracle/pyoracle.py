"""TEST INFRASTRUCTURE — ctypes access to the two CPU checkers built by oracle/Makefile.

    Oracle("port")       -> oracle/liblcg_oracle.so      (our C restatement, lcg_oracle.c)
    Oracle("reference")  -> oracle/_ref/liblcg_ref.so    (the unmodified reference CPU/OpenMP library + ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
Nothing here reads /root/reference at run time: the reference library is prebuilt and travels with the repo.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liblcg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "liblcg_ref.so")

# solver ids (reference util.h:32-64, 187-221)
LCG_CG, LCG_PCG, LCG_CGS, LCG_BICGSTAB, LCG_BICGSTAB2, LCG_PG, LCG_SPG = range(7)
CLCG_BICG, CLCG_BICG_SYM, CLCG_CGS, CLCG_BICGSTAB, CLCG_TFQMR, CLCG_PCG, CLCG_PBICG = range(7)


class LcgPara(C.Structure):
    """lcg_para, reference util.h:95-148 (64 bytes, offsets 0/8/16/24/32/40/48/56)."""
    _fields_ = [("max_iterations", C.c_int), ("epsilon", C.c_double), ("abs_diff", C.c_int),
                ("restart_epsilon", C.c_double), ("step", C.c_double), ("sigma", C.c_double),
                ("beta", C.c_double), ("maxi_m", C.c_int)]


class ClcgPara(C.Structure):
    """clcg_para, reference util.h:247-273 (24 bytes)."""
    _fields_ = [("max_iterations", C.c_int), ("epsilon", C.c_double), ("abs_diff", C.c_int)]


def default_para(**kw) -> LcgPara:
    p = LcgPara(0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10)  # defparam, util.h:153
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_cpara(**kw) -> ClcgPara:
    p = ClcgPara(0, 1e-6, 0)  # defparam2, util.h:278
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@dataclass
class SolveResult:
    ret: int
    iters: int          # k passed to the last progress call
    calls: int          # number of progress calls
    residual: float     # residual passed to the last progress call
    seconds: float      # wall time inside the solver call
    x: np.ndarray
    history: np.ndarray


def build(verbose: bool = False) -> None:
    """(Re)build the checkers.  The reference part is skipped by the Makefile when /root/reference is absent."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            path, self.pfx = PORT_SO, "lcgoracle_"
        elif kind == "reference":
            path, self.pfx = REF_SO, "lcgref_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            if kind == "port":
                build()
            else:
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._solve = getattr(self.lib, self.pfx + "solve")
        self._solve.restype = C.c_int
        self._csolve = getattr(self.lib, self.pfx + "csolve")
        self._csolve.restype = C.c_int
        getattr(self.lib, self.pfx + "num_threads").restype = C.c_int

    def num_threads(self) -> int:
        return int(getattr(self.lib, self.pfx + "num_threads")())

    def set_num_threads(self, n: int) -> None:
        """OpenMP threads of the CPU arm, independent of an inherited OMP_NUM_THREADS (torchrun exports 1)."""
        getattr(self.lib, self.pfx + "set_num_threads")(C.c_int(int(n)))

    def set_time(self, t: int) -> None:
        """Pin the seed clcg_vecrnd() derives from time(0) (reference lcg_complex.cpp:118-127)."""
        getattr(self.lib, self.pfx + "set_time")(C.c_long(t))

    def set_summation(self, tree: bool) -> None:
        """Port only: pairwise (tree) instead of the reference's left-to-right inner products — the order class of the GPU's
        reductions.  Never used for pinning; see lcgoracle_set_summation in lcg_oracle.c."""
        if self.kind != "port":
            raise ValueError("the unmodified reference sums left to right")
        self.lib.lcgoracle_set_summation(C.c_int(1 if tree else 0))

    def vecrnd(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.complex128)
        getattr(self.lib, self.pfx + "vecrnd")(_p(out, C.c_double), C.c_int(n))
        return out

    def spmv(self, A, x):
        y = np.empty_like(x)
        if np.iscomplexobj(A["val"]):
            raise ValueError("use cspmv")
        getattr(self.lib, self.pfx + "spmv")(C.c_int(A["n"]), _p(A["row_ptr"], C.c_int), _p(A["col"], C.c_int),
                                              _p(A["val"], C.c_double), _p(x, C.c_double), _p(y, C.c_double))
        return y

    def cspmv(self, A, x, transpose=False, conjugate=False):
        y = np.empty_like(x)
        getattr(self.lib, self.pfx + "cspmv")(C.c_int(A["n"]), _p(A["row_ptr"], C.c_int), _p(A["col"], C.c_int),
                                               _p(A["val"], C.c_double), _p(x, C.c_double), _p(y, C.c_double),
                                               C.c_int(int(transpose)), C.c_int(int(conjugate)))
        return y

    def solve(self, solver_id, A, b, x0=None, para=None, low=None, hig=None, diag=None,
              progress=True, stop_at=-1, hist_cap=0) -> SolveResult:
        """Real solve.  A = dict(n,row_ptr,col,val).  PCG uses Jacobi z = r/diag (diag required)."""
        n = A["n"]
        x = np.zeros(n, dtype=np.float64) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        b = np.ascontiguousarray(b, dtype=np.float64)
        hist = np.zeros(max(hist_cap, 1), dtype=np.float64)
        out = (C.c_int * 2)()
        dout = (C.c_double * 2)()
        para = para if para is not None else default_para()
        ret = self._solve(C.c_int(solver_id), C.c_int(n), _p(A["row_ptr"], C.c_int), _p(A["col"], C.c_int),
                          _p(A["val"], C.c_double), _p(x, C.c_double), _p(b, C.c_double),
                          _p(low, C.c_double), _p(hig, C.c_double), _p(diag, C.c_double), C.byref(para),
                          C.c_int(int(progress)), C.c_int(stop_at), _p(hist, C.c_double), C.c_int(hist_cap), out, dout)
        return SolveResult(ret, out[0], out[1], dout[0], dout[1], x, hist[:min(out[1], hist_cap)].copy())

    def csolve(self, solver_id, A, b, x0=None, para=None, diag=None, progress=True, stop_at=-1, hist_cap=0) -> SolveResult:
        """Complex solve.  `diag` (complex Jacobi diagonal) is only understood by the port (CLCG_PCG)."""
        n = A["n"]
        x = np.zeros(n, dtype=np.complex128) if x0 is None else np.array(x0, dtype=np.complex128, copy=True)
        b = np.ascontiguousarray(b, dtype=np.complex128)
        val = np.ascontiguousarray(A["val"], dtype=np.complex128)
        hist = np.zeros(max(hist_cap, 1), dtype=np.float64)
        out = (C.c_int * 2)()
        dout = (C.c_double * 2)()
        para = para if para is not None else default_cpara()
        args = [C.c_int(solver_id), C.c_int(n), _p(A["row_ptr"], C.c_int), _p(A["col"], C.c_int),
                _p(val, C.c_double), _p(x, C.c_double), _p(b, C.c_double)]
        if self.kind == "port":
            args.append(_p(diag, C.c_double))
        elif solver_id == CLCG_PCG:
            raise ValueError("the reference has no buildable CPU complex PCG (Eigen only)")
        args += [C.byref(para), C.c_int(int(progress)), C.c_int(stop_at), _p(hist, C.c_double), C.c_int(hist_cap), out, dout]
        ret = self._csolve(*args)
        return SolveResult(ret, out[0], out[1], dout[0], dout[1], x, hist[:min(out[1], hist_cap)].copy())


def _lower_coo(A):
    """Row-sorted COO of the full matrix A = dict(n, row_ptr, col, val)."""
    rows = np.repeat(np.arange(A["n"], dtype=np.int32), np.diff(A["row_ptr"])).astype(np.int32)
    return rows, np.ascontiguousarray(A["col"], dtype=np.int32)


def ref_ic0_half(A):
    """lcg_incomplete_Cholesky_half_coo (reference preconditioner.cpp:33-160) on a real CSR system -> (row, col, val) of L (COO)."""
    lib = C.CDLL(REF_SO)
    lib.lcgref_ic0_half.restype = C.c_int
    rows, cols = _lower_coo(A)
    val = np.ascontiguousarray(A["val"], dtype=np.float64)
    lnz = lib.lcgref_ic0_half(_p(rows, C.c_int), _p(cols, C.c_int), _p(val, C.c_double), C.c_int(A["n"]), C.c_int(len(cols)), None, None, None)
    ir, ic, iv = np.empty(lnz, np.int32), np.empty(lnz, np.int32), np.empty(lnz, np.float64)
    lib.lcgref_ic0_half(_p(rows, C.c_int), _p(cols, C.c_int), _p(val, C.c_double), C.c_int(A["n"]), C.c_int(len(cols)), _p(ir, C.c_int), _p(ic, C.c_int), _p(iv, C.c_double))
    return ir, ic, iv


def ref_cic0_half(A, single=False):
    """clcg_incomplete_Cholesky_cuda_half (reference preconditioner_cuda.cu:40-270; host code) on a complex CSR system."""
    lib = C.CDLL(REF_CUDA_SO)
    lib.lcgrefcuda_cic0_half.restype = C.c_int
    rows, cols = _lower_coo(A)
    dt = np.complex64 if single else np.complex128
    val = np.ascontiguousarray(A["val"], dtype=dt)
    args = [C.c_int(1 if single else 0), _p(rows, C.c_int), _p(cols, C.c_int), val.ctypes.data_as(C.c_void_p), C.c_int(A["n"]), C.c_int(len(cols))]
    lnz = lib.lcgrefcuda_cic0_half(*args, None, None, None)
    ir, ic, iv = np.empty(lnz, np.int32), np.empty(lnz, np.int32), np.empty(lnz, dt)
    lib.lcgrefcuda_cic0_half(*args, _p(ir, C.c_int), _p(ic, C.c_int), iv.ctypes.data_as(C.c_void_p))
    return ir, ic, iv


def ref_pcg_ic0(A, b, para=None, hist_cap=0):
    """The reference's PCG (lcg_solver_preconditioned) with M = L L^T from its own IC(0) and its own COO triangular solves as the Mx
    callback.  Returns (SolveResult, z_probe) with z_probe = M^-1 b."""
    lib = C.CDLL(REF_SO)
    lib.lcgref_pcg_ic0.restype = C.c_int
    n = A["n"]
    x = np.zeros(n)
    b = np.ascontiguousarray(b, dtype=np.float64)
    hist = np.zeros(max(hist_cap, 1))
    out, dout = (C.c_int * 2)(), (C.c_double * 2)()
    zp = np.empty(n)
    para = para if para is not None else default_para()
    ret = lib.lcgref_pcg_ic0(C.c_int(n), _p(A["row_ptr"], C.c_int), _p(A["col"], C.c_int), _p(A["val"], C.c_double), _p(x, C.c_double), _p(b, C.c_double),
                             C.byref(para), _p(hist, C.c_double), C.c_int(hist_cap), out, dout, _p(zp, C.c_double))
    return SolveResult(ret, out[0], out[1], dout[0], dout[1], x, hist[:min(out[1], hist_cap)].copy()), zp


REF_CUDA_SO = os.path.join(HERE, "_ref", "liblcg_ref_cuda.so")


def have_reference_cuda() -> bool:
    return os.path.exists(REF_CUDA_SO)


class RefCuda:
    """The UNMODIFIED reference CUDA solvers (lcg_cuda.cu: cuBLAS host loop + the caller's cusparseSpMV callback) behind
    oracle/ref_cuda_shim.cu — a GPU baseline for bench.py.  CSR arrays: device pointers; m, b: host numpy arrays."""
    SOLVERS = {"CG": 0, "PCG": 1, "CGS": 2}

    def __init__(self):
        self.lib = C.CDLL(REF_CUDA_SO)
        self.lib.lcgrefcuda_solve.restype = C.c_int

    def solve(self, solver: str, n: int, nnz: int, d_rp: int, d_ci: int, d_val: int, m: np.ndarray, b: np.ndarray,
              epsilon: float, max_iterations: int, with_progress: bool = False):
        secs, its = C.c_double(), C.c_int()
        ret = self.lib.lcgrefcuda_solve(C.c_int(self.SOLVERS[solver]), C.c_int(n), C.c_int(nnz), C.c_void_p(d_rp), C.c_void_p(d_ci), C.c_void_p(d_val),
                                        _p(m, C.c_double), _p(b, C.c_double), C.c_double(epsilon), C.c_int(max_iterations),
                                        C.c_int(int(with_progress)), C.byref(secs), C.byref(its))
        return ret, secs.value, its.value

    CSOLVERS = {"BICG": 0, "BICG_SYM": 1, "PCG": 5}

    def csolve(self, solver: str, n: int, nnz: int, d_rp: int, d_ci: int, d_val: int, m: np.ndarray, b: np.ndarray,
               epsilon: float, max_iterations: int, abs_diff: int = 0, hist_cap: int = 0):
        """clcg_solver_cuda / clcg_solver_preconditioned_cuda (clcg_cuda.cu:86-559), unmodified.  d_val: device pointer to nnz
        cuDoubleComplex; m (in/out), b: host complex128.  Returns (ret, seconds, last k, residual history)."""
        self.lib.lcgrefcuda_csolve.restype = C.c_int
        secs, its, calls = C.c_double(), C.c_int(), C.c_int()
        hist = np.zeros(max(hist_cap, 1), dtype=np.float64)
        ret = self.lib.lcgrefcuda_csolve(C.c_int(self.CSOLVERS[solver]), C.c_int(n), C.c_int(nnz), C.c_void_p(d_rp), C.c_void_p(d_ci), C.c_void_p(d_val),
                                         _p(m, C.c_double), _p(b, C.c_double), C.c_double(epsilon), C.c_int(max_iterations), C.c_int(abs_diff),
                                         _p(hist, C.c_double), C.c_int(hist_cap), C.byref(secs), C.byref(its), C.byref(calls))
        return ret, secs.value, its.value, hist[:min(calls.value, hist_cap)].copy()

    def csolvef(self, solver: str, n: int, nnz: int, d_rp: int, d_ci: int, d_val: int, m: np.ndarray, b: np.ndarray,
                epsilon: float, max_iterations: int, abs_diff: int = 0, hist_cap: int = 0):
        """The cuComplex overloads (clcg_cudaf.cu:86-558), unmodified.  d_val: device pointer to nnz cuComplex; m (in/out), b: host
        complex64.  Returns (ret, seconds, last k, residual history)."""
        assert m.dtype == np.complex64 and b.dtype == np.complex64
        self.lib.lcgrefcuda_csolvef.restype = C.c_int
        secs, its, calls = C.c_double(), C.c_int(), C.c_int()
        hist = np.zeros(max(hist_cap, 1), dtype=np.float64)
        ret = self.lib.lcgrefcuda_csolvef(C.c_int(self.CSOLVERS[solver]), C.c_int(n), C.c_int(nnz), C.c_void_p(d_rp), C.c_void_p(d_ci), C.c_void_p(d_val),
                                          m.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_double(epsilon), C.c_int(max_iterations),
                                          C.c_int(abs_diff), _p(hist, C.c_double), C.c_int(hist_cap), C.byref(secs), C.byref(its), C.byref(calls))
        return ret, secs.value, its.value, hist[:min(calls.value, hist_cap)].copy()


_KINDS = {"7pt": 0, "27pt": 1, "7pt_cd": 2}


def gen_system(kind: str, g: int, row0: int = 0, row1: int | None = None):
    """Host (OpenMP) generator of the SURVEY §8(d) stencil systems — lcgoracle_gen_stencil in lcg_oracle.c.
    Returns dict(n, nnz, row_ptr, col, val, b) for rows [row0,row1) with GLOBAL column indices."""
    if not os.path.exists(PORT_SO):
        build()
    lib = C.CDLL(PORT_SO)
    fn = lib.lcgoracle_gen_stencil
    fn.restype = C.c_longlong
    n = g ** 3
    row1 = n if row1 is None else row1
    rows = row1 - row0
    rp = np.empty(rows + 1, dtype=np.int32)
    k = C.c_int(_KINDS[kind])
    nnz = fn(k, C.c_int(g), C.c_longlong(row0), C.c_longlong(row1), _p(rp, C.c_int), None, None, None)
    ci = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    b = np.empty(rows, dtype=np.float64)
    fn(k, C.c_int(g), C.c_longlong(row0), C.c_longlong(row1), _p(rp, C.c_int), _p(ci, C.c_int), _p(val, C.c_double), _p(b, C.c_double))
    return dict(n=rows, nnz=int(nnz), row_ptr=rp, col=ci, val=val, b=b)
