"""Host-side mirror of liblcg's CUDA solver interface on top of the C ABI (include/lcgb200.h).

Names, argument meaning and return codes follow the reference (lcg_cuda.h:81-131, clcg_cuda.h:81-105):
`lcg_solver_cuda`, `lcg_solver_preconditioned_cuda`, `lcg_solver_constrained_cuda`, `clcg_solver_cuda`,
`clcg_solver_preconditioned_cuda` take HOST arrays m (in/out) and B, a parameter block and callbacks.
The matrix lives in a `CsrOperator` (the built-in fused operator); pass `CSR_AX` / `JACOBI_MX` as the Ax / Mx
callbacks and the operator as `instance`, exactly like a C++ caller passes lcgb200_csr_ax and the handle.

PyTorch is only used by callers for device memory and streams; nothing here imports it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import LcgPara, ClcgPara, Info, PROGRESS, CPROGRESS

# solver ids (reference util.h:32-64, 187-221)
LCG_CG, LCG_PCG, LCG_CGS, LCG_BICGSTAB, LCG_BICGSTAB2, LCG_PG, LCG_SPG = range(7)
CLCG_BICG, CLCG_BICG_SYM, CLCG_CGS, CLCG_BICGSTAB, CLCG_TFQMR, CLCG_PCG, CLCG_PBICG = range(7)

# return codes (reference util.h:69-90)
LCG_CONVERGENCE, LCG_STOP, LCG_ALREADY_OPTIMIZIED = 0, 1, 2
LCG_UNKNOWN_ERROR, LCG_INVILAD_VARIABLE_SIZE, LCG_INVILAD_MAX_ITERATIONS, LCG_INVILAD_EPSILON = -1024, -1023, -1022, -1021
LCG_INVILAD_RESTART_EPSILON, LCG_REACHED_MAX_ITERATIONS, LCG_NULL_PRECONDITION_MATRIX, LCG_NAN_VALUE = -1020, -1019, -1018, -1017
LCG_INVALID_POINTER, LCG_INVALID_LAMBDA, LCG_INVALID_SIGMA, LCG_INVALID_BETA, LCG_INVALID_MAXIM, LCG_SIZE_NOT_MATCH = -1016, -1015, -1014, -1013, -1012, -1011
CLCG_REACHED_MAX_ITERATIONS, CLCG_NAN_VALUE, CLCG_INVALID_POINTER, CLCG_SIZE_NOT_MATCH, CLCG_UNKNOWN_SOLVER = -1020, -1019, -1018, -1017, -1016

REAL, COMPLEX, COMPLEX_FLOAT = 0, 1, 2
HOST, DEVICE = 0, 1
CSR_TRANSPOSE, CSR_JACOBI, CSR_COMPRESS, CSR_IC0 = 1, 2, 4, 8
VEC_DEVICE, USE_JACOBI, USE_IC0 = 1, 2, 4


class _Sentinel:
    def __init__(self, symbol):
        self.symbol = symbol

    @property
    def address(self):
        return _lib.fn_addr(self.symbol)


CSR_AX = _Sentinel("lcgb200_csr_ax")
JACOBI_MX = _Sentinel("lcgb200_jacobi_mx")
CSR_CAX = _Sentinel("lcgb200_csr_cax")
JACOBI_CMX = _Sentinel("lcgb200_jacobi_cmx")
IC0_MX = _Sentinel("lcgb200_ic0_mx")      # built-in IC(0) preconditioner (operator created with ic0=True)
IC0_CMX = _Sentinel("lcgb200_ic0_cmx")


def lcg_default_parameters(**kw) -> LcgPara:
    """defparam (reference util.h:153)."""
    p = LcgPara(0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def clcg_default_parameters(**kw) -> ClcgPara:
    """defparam2 (reference util.h:278)."""
    p = ClcgPara(0, 1e-6, 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def last_error() -> str:
    return (_lib.load().lcgb200_last_error() or b"").decode()


def set_shadow_seed(seed: int) -> None:
    _lib.load().lcgb200_set_shadow_seed(seed)


def set_complex_residual_mode(mode: int) -> None:
    _lib.load().lcgb200_set_complex_residual_mode(mode)


def set_poll_interval(n: int) -> None:
    _lib.load().lcgb200_set_poll_interval(n)


def set_fused_small(on: bool) -> None:
    _lib.load().lcgb200_set_fused_small(1 if on else 0)


def set_reference_order(on: bool) -> None:
    """Reference-order arithmetic (lcgb200_set_reference_order): later solves are bit-identical to the reference's CPU solvers."""
    _lib.load().lcgb200_set_reference_order(1 if on else 0)


def set_profile(on: bool) -> None:
    _lib.load().lcgb200_set_profile(1 if on else 0)


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class CsrOperator:
    """Built-in CSR operator handle (lcgb200_csr_t).  Arrays may be numpy (host) or torch CUDA tensors (device)."""

    def __init__(self, row_ptr, col, val, n_cols=None, transpose=False, jacobi=False, compress=False, ic0=False):
        lib = _lib.load()
        on_dev = hasattr(val, "data_ptr")
        if on_dev:
            cx = val.is_complex() or (val.dim() == 2 and val.shape[-1] == 2)
            n = row_ptr.numel() - 1
            nnz = col.numel()
            single = str(val.dtype) == "torch.complex64"
            # the C ABI takes raw pointers: anything but contiguous CUDA int32 / float64 / complex128 would be reinterpreted silently
            for name, t, ok in (("row_ptr", row_ptr, ("torch.int32",)), ("col", col, ("torch.int32",)),
                                ("val", val, ("torch.complex128", "torch.complex64") if val.is_complex() else ("torch.float64",))):
                if not (t.is_cuda and t.is_contiguous() and str(t.dtype) in ok):
                    raise TypeError(f"CsrOperator: {name} must be a contiguous CUDA tensor of dtype {ok[0]} (got {t.dtype}, cuda={t.is_cuda}, contiguous={t.is_contiguous()})")
        else:
            row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
            col = np.ascontiguousarray(col, dtype=np.int32)
            cx = np.iscomplexobj(val)
            single = cx and np.asarray(val).dtype == np.complex64   # cuComplex storage: the clcg_cudaf.h entry points
            val = np.ascontiguousarray(val, dtype=(np.complex64 if single else np.complex128) if cx else np.float64)
            n = len(row_ptr) - 1
            nnz = len(col)
        self.n, self.nnz, self.complex, self.single = n, nnz, bool(cx), bool(cx and single)
        self.n_cols = n if n_cols is None else n_cols
        flags = (CSR_TRANSPOSE if transpose else 0) | (CSR_JACOBI if jacobi else 0) | (CSR_COMPRESS if compress else 0) | (CSR_IC0 if ic0 else 0)
        h = C.c_void_p()
        rc = lib.lcgb200_csr_create_rect(C.byref(h), n, self.n_cols, nnz, _ptr(row_ptr), _ptr(col), _ptr(val),
                                         (COMPLEX_FLOAT if self.single else COMPLEX) if cx else REAL, DEVICE if on_dev else HOST, flags)
        if rc != 0:
            raise RuntimeError(f"lcgb200_csr_create failed ({rc}): {last_error()}")
        self.handle = h
        self._keep = None

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().lcgb200_csr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        v = [C.c_int() for _ in range(5)]
        _lib.load().lcgb200_csr_info(self.handle, *[C.byref(x) for x in v])
        return dict(zip(("n_rows", "n_cols", "nnz", "n_tiles", "lanes_per_row"), (x.value for x in v)))

    def format(self):
        """dict(compressed, n_values, n_offsets, stream_bytes): what the SpMV streams (lcgb200_csr_format)."""
        c, nv, no, sb = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        _lib.load().lcgb200_csr_format(self.handle, C.byref(c), C.byref(nv), C.byref(no), C.byref(sb))
        return dict(compressed=bool(c.value), level=c.value, n_values=nv.value, n_offsets=no.value, stream_bytes=int(sb.value))

    def pattern_kernel(self):
        """dict(kernel, stride, n_patterns): which kernel walks the row patterns of a level-2 copy (lcgb200_csr_pattern_kernel):
        "none", "chains" (k_spmv_pat), "box" (k_spmv_pat_box) or "march" (k_spmv_pat_march)."""
        k, s, p = C.c_int(), C.c_int(), C.c_int()
        _lib.load().lcgb200_csr_pattern_kernel(self.handle, C.byref(k), C.byref(s), C.byref(p))
        return dict(kernel=("none", "chains", "box", "march")[k.value], stride=s.value, n_patterns=p.value)

    def spmv_bytes(self) -> int:
        return int(_lib.load().lcgb200_csr_spmv_bytes(self.handle))

    def diagonal(self) -> np.ndarray:
        out = np.empty(self.n, dtype=(np.complex64 if getattr(self, "single", False) else np.complex128) if self.complex else np.float64)
        rc = _lib.load().lcgb200_csr_get_diagonal(self.handle, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"get_diagonal failed ({rc})")
        return out

    def ic0_factor(self):
        """The IC(0) factor L of an operator created with ic0=True: dict(row_ptr, col, val, levels_lower, levels_upper)."""
        lib = _lib.load()
        lnz, ll, lu = C.c_int(), C.c_int(), C.c_int()
        rc = lib.lcgb200_csr_get_ic0(self.handle, C.byref(lnz), None, None, None, C.byref(ll), C.byref(lu))
        if rc != 0:
            raise RuntimeError(f"lcgb200_csr_get_ic0 failed ({rc}): {last_error()}")
        rp = np.empty(self.n + 1, dtype=np.int32)
        ci = np.empty(lnz.value, dtype=np.int32)
        val = np.empty(lnz.value, dtype=(np.complex64 if getattr(self, "single", False) else np.complex128) if self.complex else np.float64)
        rc = lib.lcgb200_csr_get_ic0(self.handle, None, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, None, None)
        if rc != 0:
            raise RuntimeError(f"lcgb200_csr_get_ic0 failed ({rc}): {last_error()}")
        return dict(row_ptr=rp, col=ci, val=val, levels_lower=ll.value, levels_upper=lu.value)

    def ic0_apply(self, r_dev, z_dev, stream=None):
        """z = (L L^T)^-1 r on device vectors (two sparse triangular solves)."""
        rc = _lib.load().lcgb200_csr_ic0_apply(self.handle, _ptr(r_dev), _ptr(z_dev), stream)
        if rc != 0:
            raise RuntimeError(f"ic0_apply failed ({rc}): {last_error()}")

    def spmv(self, x_dev, y_dev, op=0, stream=None):
        rc = _lib.load().lcgb200_csr_spmv(self.handle, _ptr(x_dev), _ptr(y_dev), op, stream)
        if rc != 0:
            raise RuntimeError(f"spmv failed ({rc}): {last_error()}")

    def spmv_dot(self, x_dev, y_dev, w_dev, dots_dev, stream=None):
        rc = _lib.load().lcgb200_csr_spmv_dot(self.handle, _ptr(x_dev), _ptr(y_dev), _ptr(w_dev), _ptr(dots_dev), stream)
        if rc != 0:
            raise RuntimeError(f"spmv_dot failed ({rc}): {last_error()}")


def read_case(path_a: str, complex_valued: bool = False):
    """lcgb200_read_case: a reference fixture (data/case_*_A) -> dict(n, nnz, rows, cols, vals, b), COO triplets row-sorted."""
    lib = _lib.load()
    n, nz = C.c_int(), C.c_int()
    rows, cols, vals, rhs = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = lib.lcgb200_read_case(path_a.encode(), COMPLEX if complex_valued else REAL, C.byref(n), C.byref(nz), C.byref(rows), C.byref(cols),
                               C.byref(vals), C.byref(rhs))
    if rc != 0:
        raise RuntimeError(f"lcgb200_read_case failed ({rc}): {last_error()}")
    vt = np.complex128 if complex_valued else np.float64
    try:
        out = dict(n=n.value, nnz=nz.value,
                   rows=np.ctypeslib.as_array(C.cast(rows, C.POINTER(C.c_int)), shape=(max(nz.value, 1),))[:nz.value].copy(),
                   cols=np.ctypeslib.as_array(C.cast(cols, C.POINTER(C.c_int)), shape=(max(nz.value, 1),))[:nz.value].copy(),
                   vals=np.frombuffer(C.string_at(vals, nz.value * np.dtype(vt).itemsize), dtype=vt).copy(),
                   b=np.frombuffer(C.string_at(rhs, n.value * np.dtype(vt).itemsize), dtype=vt).copy())
    finally:
        for p in (rows, cols, vals, rhs):
            lib.lcgb200_free_host(p)
    return out


def operator_from_coo(n, rows, cols, vals, transpose=False, jacobi=False) -> "CsrOperator":
    """lcgb200_csr_create_from_coo: row-sorted host COO triplets -> built-in operator (COO -> CSR compressed on the device)."""
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    cx = np.iscomplexobj(vals)
    vals = np.ascontiguousarray(vals, dtype=np.complex128 if cx else np.float64)
    h = C.c_void_p()
    flags = (CSR_TRANSPOSE if transpose else 0) | (CSR_JACOBI if jacobi else 0)
    rc = lib.lcgb200_csr_create_from_coo(C.byref(h), n, len(rows), rows.ctypes.data, cols.ctypes.data, vals.ctypes.data, COMPLEX if cx else REAL, flags)
    if rc != 0:
        raise RuntimeError(f"lcgb200_csr_create_from_coo failed ({rc}): {last_error()}")
    op = CsrOperator.__new__(CsrOperator)
    op.n, op.nnz, op.complex, op.n_cols, op.handle, op._keep, op.single = n, len(rows), bool(cx), n, h, None, False
    return op


@dataclass
class Result:
    ret: int
    iterations: int
    residual: float
    info: Info


def _wrap_progress(Pfp, complex_):
    """Python callable (instance, m_dev_ptr, converge, param, n, nz, k) -> int  =>  C callback."""
    if Pfp is None:
        return None, None
    proto = CPROGRESS if complex_ else PROGRESS

    def tramp(instance, m_dev, converge, param, n, nz, k):
        return int(Pfp(instance, m_dev, converge, param.contents, n, nz, k) or 0)

    cb = proto(tramp)
    return cb, C.cast(cb, C.c_void_p)


def _afp_addr(Afp):
    if Afp is None:
        return None
    if isinstance(Afp, _Sentinel):
        return Afp.address
    if isinstance(Afp, int):
        return Afp
    return C.cast(Afp, C.c_void_p).value


def _instance(instance):
    return instance.handle if isinstance(instance, CsrOperator) else instance


# ---------------------------------------------------------------------------------- reference-shaped calls
def lcg_solver_cuda(Afp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None, solver_id=LCG_CG) -> int:
    """lcg_solver_cuda (reference lcg_cuda.h:81-83).  m (in/out) and B are host float64 arrays."""
    cb, cbp = _wrap_progress(Pfp, False)
    return _lib.load().lcgb200_solver_cuda(_afp_addr(Afp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                           C.byref(param) if param is not None else None, _instance(instance),
                                           cub_handle, cus_handle, solver_id)


def lcg_solver_preconditioned_cuda(Afp, Mfp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None,
                                   solver_id=LCG_PCG) -> int:
    """lcg_solver_preconditioned_cuda (reference lcg_cuda.h:104-106)."""
    cb, cbp = _wrap_progress(Pfp, False)
    return _lib.load().lcgb200_solver_preconditioned_cuda(_afp_addr(Afp), _afp_addr(Mfp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                                          C.byref(param) if param is not None else None, _instance(instance),
                                                          cub_handle, cus_handle, solver_id)


def lcg_solver_constrained_cuda(Afp, Pfp, m, B, low, hig, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None,
                                solver_id=LCG_PG) -> int:
    """lcg_solver_constrained_cuda (reference lcg_cuda.h:129-131)."""
    cb, cbp = _wrap_progress(Pfp, False)
    return _lib.load().lcgb200_solver_constrained_cuda(_afp_addr(Afp), cbp, _ptr(m), _ptr(B), _ptr(low), _ptr(hig), n_size, nz_size,
                                                       C.byref(param) if param is not None else None, _instance(instance),
                                                       cub_handle, cus_handle, solver_id)


def clcg_solver_cuda(Afp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None, solver_id=CLCG_BICG) -> int:
    """clcg_solver_cuda (reference clcg_cuda.h:81-83).  m, B: host complex128 arrays."""
    cb, cbp = _wrap_progress(Pfp, True)
    return _lib.load().lcgb200_csolver_cuda(_afp_addr(Afp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                            C.byref(param) if param is not None else None, _instance(instance),
                                            cub_handle, cus_handle, solver_id)


def clcg_solver_preconditioned_cuda(Afp, Mfp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None,
                                    solver_id=CLCG_PCG) -> int:
    """clcg_solver_preconditioned_cuda (reference clcg_cuda.h:103-105)."""
    cb, cbp = _wrap_progress(Pfp, True)
    return _lib.load().lcgb200_csolver_preconditioned_cuda(_afp_addr(Afp), _afp_addr(Mfp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                                           C.byref(param) if param is not None else None, _instance(instance),
                                                           cub_handle, cus_handle, solver_id)


def _wrap_progress_f(Pfp):
    if Pfp is None:
        return None, None

    def tramp(instance, m_dev, converge, param, n, nz, k):
        return int(Pfp(instance, m_dev, converge, param.contents, n, nz, k) or 0)

    cb = _lib.CPROGRESSF(tramp)
    return cb, C.cast(cb, C.c_void_p)


def clcg_solver_cudaf(Afp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None, solver_id=CLCG_BICG) -> int:
    """The cuComplex overload of clcg_solver_cuda (reference clcg_cudaf.h:81-83).  m, B: host complex64 arrays."""
    cb, cbp = _wrap_progress_f(Pfp)
    return _lib.load().lcgb200_csolver_cudaf(_afp_addr(Afp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                             C.byref(param) if param is not None else None, _instance(instance),
                                             cub_handle, cus_handle, solver_id)


def clcg_solver_preconditioned_cudaf(Afp, Mfp, Pfp, m, B, n_size, nz_size, param, instance, cub_handle=None, cus_handle=None,
                                     solver_id=CLCG_PCG) -> int:
    """The cuComplex overload of clcg_solver_preconditioned_cuda (reference clcg_cudaf.h:103-105)."""
    cb, cbp = _wrap_progress_f(Pfp)
    return _lib.load().lcgb200_csolver_preconditioned_cudaf(_afp_addr(Afp), _afp_addr(Mfp), cbp, _ptr(m), _ptr(B), n_size, nz_size,
                                                            C.byref(param) if param is not None else None, _instance(instance),
                                                            cub_handle, cus_handle, solver_id)


# ------------------------------------------------------------------ the reference's HOST-callback API (lcg.h, clcg.h)
CSR_AX_HOST = _Sentinel("lcgb200_csr_ax_host")
JACOBI_MX_HOST = _Sentinel("lcgb200_jacobi_mx_host")
CSR_CAX_HOST = _Sentinel("lcgb200_csr_cax_host")


def _host_ax(fn, n, complex_):
    """Python callable (x: ndarray[, layout, conjugate]) -> ndarray  =>  C host callback."""
    if fn is None or isinstance(fn, _Sentinel):
        return fn, _afp_addr(fn)
    k = 2 if complex_ else 1

    if complex_:
        def tramp(instance, x, y, nn, layout, conj):
            xv = np.ctypeslib.as_array(x, shape=(nn * k,)).view(np.complex128)
            np.ctypeslib.as_array(y, shape=(nn * k,)).view(np.complex128)[:] = fn(xv, layout, conj)
        cb = _lib.CAXFUNC_HOST(tramp)
    else:
        def tramp(instance, x, y, nn):
            np.ctypeslib.as_array(y, shape=(nn,))[:] = fn(np.ctypeslib.as_array(x, shape=(nn,)))
        cb = _lib.AXFUNC_HOST(tramp)
    return cb, C.cast(cb, C.c_void_p).value


def _host_progress(fn, complex_):
    if fn is None:
        return None, None
    k = 2 if complex_ else 1

    def tramp(instance, m, converge, param, n, it):
        mv = np.ctypeslib.as_array(m, shape=(n * k,))
        return int(fn(mv.view(np.complex128) if complex_ else mv, converge, param.contents, n, it) or 0)

    cb = (_lib.CPROGRESS_HOST if complex_ else _lib.PROGRESS_HOST)(tramp)
    return cb, C.cast(cb, C.c_void_p).value


def lcg_solver(Afp, Pfp, m, B, n_size, param, instance, solver_id=LCG_CGS) -> int:
    """lcg_solver (reference lcg.h:71-72).  Afp: CSR_AX_HOST (instance = CsrOperator) or a Python callable x -> A x run on the host."""
    a, ap = _host_ax(Afp, n_size, False)
    p, pp = _host_progress(Pfp, False)
    return _lib.load().lcgb200_solver(ap, pp, _ptr(m), _ptr(B), n_size, C.byref(param) if param is not None else None, _instance(instance), solver_id)


def lcg_solver_preconditioned(Afp, Mfp, Pfp, m, B, n_size, param, instance, solver_id=LCG_PCG) -> int:
    """lcg_solver_preconditioned (reference lcg.h:90-91)."""
    a, ap = _host_ax(Afp, n_size, False)
    mm, mp = _host_ax(Mfp, n_size, False)
    p, pp = _host_progress(Pfp, False)
    return _lib.load().lcgb200_solver_preconditioned(ap, mp, pp, _ptr(m), _ptr(B), n_size, C.byref(param) if param is not None else None,
                                                     _instance(instance), solver_id)


def lcg_solver_constrained(Afp, Pfp, m, B, low, hig, n_size, param, instance, solver_id=LCG_PG) -> int:
    """lcg_solver_constrained (reference lcg.h:111-113)."""
    a, ap = _host_ax(Afp, n_size, False)
    p, pp = _host_progress(Pfp, False)
    return _lib.load().lcgb200_solver_constrained(ap, pp, _ptr(m), _ptr(B), _ptr(low), _ptr(hig), n_size, C.byref(param) if param is not None else None,
                                                  _instance(instance), solver_id)


def clcg_solver(Afp, Pfp, m, B, n_size, param, instance, solver_id=CLCG_BICG) -> int:
    """clcg_solver (reference clcg.h:74-76).  Afp: CSR_CAX_HOST or a Python callable (x, layout, conjugate) -> op(A) x."""
    a, ap = _host_ax(Afp, n_size, True)
    p, pp = _host_progress(Pfp, True)
    return _lib.load().lcgb200_csolver(ap, pp, _ptr(m), _ptr(B), n_size, C.byref(param) if param is not None else None, _instance(instance), solver_id)


# ------------------------------------------------------------------------------------- handle-shaped calls
def ic0_factor_host(row_ptr, col, val):
    """lcgb200_ic0_factor_host: IC(0) of the lower triangle given as CSR (diagonal last in every row), on the host, in place."""
    val = np.ascontiguousarray(val)
    vt = {np.dtype(np.float64): REAL, np.dtype(np.complex128): COMPLEX, np.dtype(np.complex64): COMPLEX_FLOAT}[val.dtype]
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(col, dtype=np.int32)
    out = val.copy()
    rc = _lib.load().lcgb200_ic0_factor_host(len(rp) - 1, rp.ctypes.data, ci.ctypes.data, out.ctypes.data, vt)
    if rc != 0:
        raise RuntimeError(f"lcgb200_ic0_factor_host failed ({rc})")
    return out


def solve(A: CsrOperator, solver_id, m, B, low=None, hig=None, param=None, Pfp=None, device=False, jacobi=False, stream=None, ic0=False) -> Result:
    """lcgb200_solve: real solvers on the built-in operator; m/B numpy (host) or CUDA tensors (device=True)."""
    cb, cbp = _wrap_progress(Pfp, False)
    info = Info()
    flags = (VEC_DEVICE if device else 0) | (USE_JACOBI if jacobi else 0) | (USE_IC0 if ic0 else 0)
    rc = _lib.load().lcgb200_solve(A.handle, solver_id, _ptr(m), _ptr(B), _ptr(low), _ptr(hig),
                                   C.byref(param) if param is not None else None, cbp, flags, stream, C.byref(info))
    return Result(rc, info.iterations, info.residual, info)


def csolve(A: CsrOperator, solver_id, m, B, param=None, Pfp=None, device=False, jacobi=False, stream=None, ic0=False) -> Result:
    """lcgb200_csolve: complex solvers on the built-in operator."""
    cb, cbp = _wrap_progress(Pfp, True)
    info = Info()
    flags = (VEC_DEVICE if device else 0) | (USE_JACOBI if jacobi else 0) | (USE_IC0 if ic0 else 0)
    rc = _lib.load().lcgb200_csolve(A.handle, solver_id, _ptr(m), _ptr(B), C.byref(param) if param is not None else None,
                                    cbp, flags, stream, C.byref(info))
    return Result(rc, info.iterations, info.residual, info)
