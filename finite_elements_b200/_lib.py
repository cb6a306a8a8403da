"""ctypes binding of libfe_b200.so (include/fe_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
The product never computes the hot path on the CPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FE_B200_LIB") or os.path.join(_HERE, "libfe_b200.so")  # env: tuning builds only

FE_OK = 0
FE_ERR_ARG, FE_ERR_CUDA, FE_ERR_NCCL = -1, -2, -3
FE_ERR_NOT_CONVERGED, FE_ERR_BREAKDOWN, FE_ERR_UNSUPPORTED = -4, -5, -6

KIND_ELAST_PSTRESS, KIND_ELAST_PSTRAIN, KIND_MAGNETIC, KIND_MASS = 0, 1, 2, 3
KIND_ELAST_TET, KIND_MASS_TET = 4, 5


class NotConverged(RuntimeError):
    """PCG reached maxit (FE_ERR_NOT_CONVERGED); .iters / .relres are set by the caller."""


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); must list every symbol declared in include/fe_b200.h
SIGNATURES = {
    "fe_version": (C.c_int, []),
    "fe_last_error": (C.c_char_p, []),
    "fe_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "fe_ctx_destroy": (C.c_int, [_vp]),
    "fe_ctx_launch_count": (_i64, [_vp]),
    "fe_elem_matrices": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fe_elem_post": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fe_source_factors": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "fe_plan_create": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _i32, _vp, _vp, C.POINTER(_vp)]),
    "fe_plan_destroy": (C.c_int, [_vp]),
    "fe_plan_nnz": (_i64, [_vp]),
    "fe_plan_n_rows": (_i32, [_vp]),
    "fe_plan_max_degree": (_i32, [_vp]),
    "fe_plan_bytes": (_i64, [_vp]),
    "fe_plan_fan_record_bytes": (_i32, [_vp]),
    "fe_plan_csr": (C.c_int, [_vp, _vp, _vp, _vp]),
    "fe_assemble": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _i32, _vp, C.c_int]),
    "fe_dirichlet_apply": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fe_scatter_add": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "fe_spmv": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32]),
    "fe_tet_elem_matrices": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fe_tet_elem_post": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fe_tet_plan_create": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, C.POINTER(_vp)]),
    "fe_tet_assemble": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _i32, _vp, _i32]),
    "fe_spmm_pair": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32]),
    "fe_cheb_step": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _i32, _i32]),
    "fe_csr_diagonal": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "fe_pcg_work_len": (_i64, [_i32, _i32]),
    "fe_pcg_cache_pattern": (C.c_int, [_vp, _vp, _vp, _i64]),
    "fe_pcg": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f64, _i32,
                         C.POINTER(_i32), C.POINTER(_f64)]),
    "fe_pcg_fixed": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32]),
    "fe_dist_unique_id": (C.c_int, [_vp]),
    "fe_dist_init": (C.c_int, [_vp, _vp, _i32, _i32]),
    "fe_dist_p2p_export": (C.c_int, [_vp, _i32, _vp]),
    "fe_dist_p2p_import": (C.c_int, [_vp, _vp]),
    "fe_dist_pcg": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32,
                              _f64, _i32, _i32, C.POINTER(_i32), C.POINTER(_f64)]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C finite_elements_b200/csrc`). finite_elements_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error():
    msg = lib.fe_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Map C-ABI status codes to the reference's exception types:
    bad argument -> ValueError; singular / non-SPD system -> NotImplementedError
    (the reference raises it from MatrixRankWarning, analysis.py:824-826)."""
    if rc == FE_OK:
        return
    msg = last_error()
    if rc == FE_ERR_ARG:
        raise ValueError(msg)
    if rc in (FE_ERR_BREAKDOWN, FE_ERR_UNSUPPORTED):
        raise NotImplementedError(msg)
    if rc == FE_ERR_NOT_CONVERGED:
        raise NotConverged(msg)
    raise RuntimeError(f"libfe_b200 error {rc}: {msg}")
