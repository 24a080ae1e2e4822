"""ctypes binding of liblcgb200.so (the C ABI declared in include/lcgb200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the symbol table,
and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LCGB200_LIB", os.path.join(HERE, "liblcgb200.so"))   # override: tuning variants only


class LcgPara(C.Structure):
    """lcgb200_para == lcg_para (reference util.h:95-148)."""
    _fields_ = [("max_iterations", C.c_int), ("epsilon", C.c_double), ("abs_diff", C.c_int),
                ("restart_epsilon", C.c_double), ("step", C.c_double), ("sigma", C.c_double),
                ("beta", C.c_double), ("maxi_m", C.c_int)]


class ClcgPara(C.Structure):
    """lcgb200_cpara == clcg_para (reference util.h:247-273)."""
    _fields_ = [("max_iterations", C.c_int), ("epsilon", C.c_double), ("abs_diff", C.c_int)]


class Info(C.Structure):
    _fields_ = [("iterations", C.c_int), ("checks", C.c_int), ("spmv_launches", C.c_int), ("kernel_launches", C.c_int),
                ("residual", C.c_double), ("device_ms", C.c_double), ("total_ms", C.c_double),
                ("spmv_ms", C.c_double), ("vec_ms", C.c_double), ("spmv_timed", C.c_int), ("vec_timed", C.c_int)]


# callback prototypes (lcg_cuda.h:45-46,61-62; clcg_cuda.h:45-46,61-62)
AXFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int)
PROGRESS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(LcgPara), C.c_int, C.c_int, C.c_int)
CAXFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int)
CPROGRESS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(ClcgPara), C.c_int, C.c_int, C.c_int)
CPROGRESSF = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.POINTER(ClcgPara), C.c_int, C.c_int, C.c_int)   # clcg_cudaf.h:61-62
# host-callback API (lcg.h:37-38,53-54; clcg.h:40-41,56-57)
AXFUNC_HOST = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)
PROGRESS_HOST = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.POINTER(LcgPara), C.c_int, C.c_int)
CAXFUNC_HOST = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int)
CPROGRESS_HOST = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_double, C.POINTER(ClcgPara), C.c_int, C.c_int)

# every symbol include/lcgb200.h declares: name -> (restype, argtypes or None)
_VP, _I, _D, _LL = C.c_void_p, C.c_int, C.c_double, C.c_longlong
SYMBOLS = {
    "lcgb200_csr_create": (_I, [C.POINTER(_VP), _I, _I, _VP, _VP, _VP, _I, _I, C.c_uint]),
    "lcgb200_csr_create_rect": (_I, [C.POINTER(_VP), _I, _I, _I, _VP, _VP, _VP, _I, _I, C.c_uint]),
    "lcgb200_csr_destroy": (_I, [_VP]),
    "lcgb200_csr_set_user": (_I, [_VP, _VP]),
    "lcgb200_csr_get_diagonal": (_I, [_VP, _VP]),
    "lcgb200_csr_get_ic0": (_I, [_VP, C.POINTER(_I), _VP, _VP, _VP, C.POINTER(_I), C.POINTER(_I)]),
    "lcgb200_csr_ic0_apply": (_I, [_VP, _VP, _VP, _VP]),
    "lcgb200_ic0_factor_host": (_I, [_I, _VP, _VP, _VP, _I]),
    "lcgb200_ic0_mx": (None, None),
    "lcgb200_ic0_cmx": (None, None),
    "lcgb200_ic0_mx_host": (None, None),
    "lcgb200_csr_spmv": (_I, [_VP, _VP, _VP, _I, _VP]),
    "lcgb200_csr_spmv_dot": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "lcgb200_csr_spmv_bytes": (_LL, [_VP]),
    "lcgb200_csr_format": (_I, [_VP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_LL)]),
    "lcgb200_csr_pattern_kernel": (_I, [_VP] + [C.POINTER(_I)] * 3),
    "lcgb200_csr_info": (_I, [_VP] + [C.POINTER(_I)] * 5),
    "lcgb200_coo2csr": (_I, [_VP, _I, _I, _VP, _VP]),
    "lcgb200_vec_elementwise": (_I, [_I, _I, _VP, _VP, _VP, _I, _VP]),
    "lcgb200_diagonal_of_csr": (_I, [_I, _VP, _VP, _VP, _I, _VP, _VP]),
    "lcgb200_set2box": (_I, [_VP, _VP, _VP, _I, _VP]),
    "lcgb200_read_case": (_I, [C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "lcgb200_free_host": (None, [_VP]),
    "lcgb200_csr_create_from_coo": (_I, [C.POINTER(_VP), _I, _I, _VP, _VP, _VP, _I, C.c_uint]),
    "lcgb200_comm_unique_id": (_I, [_VP, _I]),
    "lcgb200_comm_create": (_I, [C.POINTER(_VP), _I, _I, _VP]),
    "lcgb200_comm_destroy": (_I, [_VP]),
    "lcgb200_comm_stats": (_I, [_VP, C.POINTER(_I), C.POINTER(_I)]),
    "lcgb200_csr_set_partition": (_I, [_VP, _VP, _LL, _I, _VP, _VP, _VP, _VP]),
    "lcgb200_csr_attach_transpose": (_I, [_VP, _VP]),
    "lcgb200_csr_set_row_offset": (_I, [_VP, _LL]),
    "lcgb200_comm_p2p_handle": (_I, [_VP, _VP, C.POINTER(_LL)]),
    "lcgb200_comm_p2p_attach": (_I, [_VP, _VP, _VP, _VP]),
    "lcgb200_comm_p2p_detach": (_I, [_VP]),
    "lcgb200_csr_ax": (None, None),
    "lcgb200_jacobi_mx": (None, None),
    "lcgb200_csr_cax": (None, None),
    "lcgb200_jacobi_cmx": (None, None),
    "lcgb200_solver_cuda": (_I, [_VP, _VP, _VP, _VP, _I, _I, C.POINTER(LcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_solver_preconditioned_cuda": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, C.POINTER(LcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_solver_constrained_cuda": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, C.POINTER(LcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_csolver_cuda": (_I, [_VP, _VP, _VP, _VP, _I, _I, C.POINTER(ClcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_csolver_preconditioned_cuda": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, C.POINTER(ClcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_csolver_cudaf": (_I, [_VP, _VP, _VP, _VP, _I, _I, C.POINTER(ClcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_csolver_preconditioned_cudaf": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, C.POINTER(ClcgPara), _VP, _VP, _VP, _I]),
    "lcgb200_csr_ax_host": (None, None),
    "lcgb200_jacobi_mx_host": (None, None),
    "lcgb200_csr_cax_host": (None, None),
    "lcgb200_solver": (_I, [_VP, _VP, _VP, _VP, _I, C.POINTER(LcgPara), _VP, _I]),
    "lcgb200_lcg": (_I, [_VP, _VP, _VP, _VP, _I, C.POINTER(LcgPara), _VP, _VP, _VP, _VP]),
    "lcgb200_lcgs": (_I, [_VP, _VP, _VP, _VP, _I, C.POINTER(LcgPara), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lcgb200_solver_preconditioned": (_I, [_VP, _VP, _VP, _VP, _VP, _I, C.POINTER(LcgPara), _VP, _I]),
    "lcgb200_solver_constrained": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, C.POINTER(LcgPara), _VP, _I]),
    "lcgb200_csolver": (_I, [_VP, _VP, _VP, _VP, _I, C.POINTER(ClcgPara), _VP, _I]),
    "lcgb200_solve": (_I, [_VP, _I, _VP, _VP, _VP, _VP, C.POINTER(LcgPara), _VP, C.c_uint, _VP, C.POINTER(Info)]),
    "lcgb200_csolve": (_I, [_VP, _I, _VP, _VP, C.POINTER(ClcgPara), _VP, C.c_uint, _VP, C.POINTER(Info)]),
    "lcgb200_set_shadow_seed": (None, [C.c_long]),
    "lcgb200_set_complex_residual_mode": (None, [_I]),
    "lcgb200_set_poll_interval": (None, [_I]),
    "lcgb200_set_profile": (None, [_I]),
    "lcgb200_set_fused_small": (None, [_I]),
    "lcgb200_set_spin_timeout_ms": (None, [_LL]),
    "lcgb200_set_graphs": (None, [_I]),
    "lcgb200_set_pdl": (None, [_I]),
    "lcgb200_set_reference_order": (None, [_I]),
    "lcgb200_set_l2_persist": (None, [_I]),
    "lcgb200_last_error": (C.c_char_p, []),
    "lcgb200_version": (_I, []),
    "lcgb200_gen_stencil": (_I, [_I, _I, _LL, _LL, _VP, _VP, _VP, _LL, C.POINTER(_LL), _VP]),
    "lcgb200_gen_rhs": (_I, [_I, _I, _LL, _LL, _VP, _VP]),
}

_lib = None


def load() -> C.CDLL:
    """Load liblcgb200.so (once) and type every exported function.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(liblcg_b200 has no CPU fallback)")
    lib = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        if args is not None:
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def fn_addr(name: str) -> int:
    """Address of an exported function (used to pass the sentinel callbacks)."""
    return C.cast(getattr(load(), name), C.c_void_p).value
