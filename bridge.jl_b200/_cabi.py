"""ctypes binding of libbridge_b200.so (include/bridge_b200.h).

This is the same binding a Julia `ccall` shim makes (julia/BridgeB200.jl); Python
is used here because no Julia runtime exists in this environment.  There is no CPU fallback: if
the shared library is missing the import fails, and without a CUDA device every compute call
raises BridgeError(BB_ERR_NODEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BB_LIB") or os.path.join(_HERE, "lib", "libbridge_b200.so")  # BB_LIB: tuning builds

BB_NPAR = 32
# status codes
OK, ERR_LENGTH, ERR_TIMEAXIS, ERR_STARTPOINT, ERR_DIM, ERR_ASSERT_M, ERR_MODEL, ERR_ARG, ERR_CUDA, \
    ERR_NOMEM, ERR_NODEVICE, ERR_UNSUPPORTED, ERR_SINGULAR, ERR_STALE, ERR_COMM, ERR_USERSRC = (
        0, -1, -2, -3, -4, -5, -6, -7, -8, -9, -10, -11, -12, -13, -14, -15)
NCCL_ID_BYTES = 128
# model ids
WIENER, OU, LINPRO, FHN_DIAG, FHN_HYPO, INTDIFF, NCLAR3, LORENZ, LANDMARKS, BOLUS, USER = range(11)
GUIDE_NUH, GUIDE_HV, GUIDE_LMMU = 1, 2, 3
ODE_R3, ODE_LYAP = 0, 1
ENS_DOUBLE_BUFFER, ENS_NO_X = 1, 2
W, X = 0, 1
CUR, PROP = 0, 1
F_LL, F_LL_PROP, F_LOGU, F_XEND, F_XEND_PROP = range(5)
RUN_STORE_X, RUN_NO_LL, RUN_SKIP_REJECTED = 1, 2, 4
ARITH_REFERENCE, ARITH_FUSED = 0, 1
PCN_AUTO, PCN_ONE_THREAD, PCN_WARP_SPECIALISED, PCN_WARP_SPECIALISED_2 = 0, 1, 2, 3
SCHEME_EULER, SCHEME_STRATONOVICH, SCHEME_HEUN, SCHEME_SRK, SCHEME_MDB = range(5)


class BridgeError(RuntimeError):
    """Raised for every non-zero bb_status; .status holds the code, the text is the reference's
    own error message where the reference has one (src/euler.jl:137,248,251; src/sde!.jl:30)."""

    def __init__(self, status: int, text: str):
        super().__init__(text)
        self.status = status


class Model(C.Structure):
    _fields_ = [("id", C.c_int32), ("d", C.c_int32), ("dprime", C.c_int32), ("reserved", C.c_int32),
                ("par", C.c_double * BB_NPAR)]


BB_NTHETA, BB_MAXD, BB_MAXSEG = 8, 4, 16
AUX_FHN_MATCHING, AUX_FHN_LINEARISED_END, AUX_BOLUS = 1, 2, 3
PRIOR_FLAT, PRIOR_GAMMA = 0, 1


class ThetaSpec(C.Structure):  # bb_theta_spec
    _fields_ = [("m", C.c_int32), ("aux_kind", C.c_int32), ("L", C.c_double * (BB_MAXD * BB_MAXD)),
                ("Sigma", C.c_double * (BB_MAXD * BB_MAXD)), ("eps", C.c_double),
                ("v", (C.c_double * BB_MAXD) * BB_MAXSEG), ("prior_kind", C.c_int32 * BB_NTHETA),
                ("prior_a", C.c_double * BB_NTHETA), ("prior_b", C.c_double * BB_NTHETA),
                ("start_sd", C.c_double), ("start_dir", C.c_double * BB_MAXD)]


class Aux(C.Structure):
    _fields_ = [("d", C.c_int32), ("is_const", C.c_int32), ("B", C.c_void_p), ("beta", C.c_void_p),
                ("a", C.c_void_p), ("a_left", C.c_void_p)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C bridge.jl_b200/csrc` "
            "(or __graft_entry__.build()); this package has no CPU implementation")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    pp = C.POINTER(C.c_void_p)
    sig = {
        "bb_strerror": (C.c_char_p, [C.c_int]),
        "bb_last_cuda_error": (C.c_char_p, []),
        "bb_abi_version": (C.c_int, []),
        "bb_ctx_create": (C.c_int, [C.c_int, pp]),
        "bb_ctx_destroy": (C.c_int, [vp]),
        "bb_ctx_synchronize": (C.c_int, [vp]),
        "bb_ctx_set_stream": (C.c_int, [vp, vp]),
        "bb_ctx_get_stream": (vp, [vp]),
        "bb_ctx_launch_count": (i64, [vp]),
        "bb_ctx_set_timing": (C.c_int, [vp, C.c_int]),
        "bb_ctx_last_kernel_ms": (dbl, [vp]),
        "bb_ctx_set_arith": (C.c_int, [vp, C.c_int]),
        "bb_ctx_set_pcn_kernel": (C.c_int, [vp, C.c_int]),
        "bb_ens_create": (C.c_int, [vp, i64, i32, i32, i32, i32, u32, pp]),
        "bb_ens_destroy": (C.c_int, [vp]),
        "bb_ens_set_chain_offset": (C.c_int, [vp, i64]),
        "bb_ens_set_grid": (C.c_int, [vp, i32, vp, i32]),
        "bb_ens_get_grid": (C.c_int, [vp, i32, vp, i32]),
        "bb_ens_set_start": (C.c_int, [vp, vp, i32, i32]),
        "bb_ens_upload": (C.c_int, [vp, C.c_int, C.c_int, i64, i64, vp]),
        "bb_ens_download": (C.c_int, [vp, C.c_int, C.c_int, i64, i64, vp]),
        "bb_ens_get_f64": (C.c_int, [vp, C.c_int, i64, i64, vp]),
        "bb_ens_set_ll": (C.c_int, [vp, i64, i64, vp]),
        "bb_ens_get_accepted": (C.c_int, [vp, i64, i64, vp]),
        "bb_ens_get_acc": (C.c_int, [vp, C.POINTER(i64)]),
        "bb_ens_reset_acc": (C.c_int, [vp]),
        "bb_ens_acc_device_ptr": (vp, [vp]),
        "bb_ens_bytes": (i64, [vp]),
        "bb_wiener_sample": (C.c_int, [vp, u64, u32]),
        "bb_euler": (C.c_int, [vp, C.POINTER(Model)]),
        "bb_guided_mdb": (C.c_int, [vp, C.POINTER(Model), pp]),
        "bb_sample_euler": (C.c_int, [vp, C.POINTER(Model), u64, u32]),
        "bb_solve_scheme": (C.c_int, [vp, C.POINTER(Model), i32]),
        "bb_guide_create": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, pp]),
        "bb_guide_create_ncd": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, pp]),
        "bb_guide_destroy": (C.c_int, [vp]),
        "bb_update_nuHC": (C.c_int, [vp, i32, i32, vp, vp, vp, dbl, vp, vp, C.POINTER(dbl)]),
        "bb_gpupdate_nuH": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp]),
        "bb_gpupdate_HV": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp]),
        "bb_backward_nuH": (C.c_int, [vp, i32, i32, i32, vp, C.POINTER(Aux), vp, vp, dbl, vp, vp, vp, vp,
                                      C.POINTER(dbl)]),
        "bb_backward_FH": (C.c_int, [vp, i32, i32, vp, C.POINTER(Aux), vp, vp, dbl, vp, vp, C.POINTER(dbl)]),
        "bb_backward_HV": (C.c_int, [vp, i32, i32, vp, C.POINTER(Aux), vp, vp, vp, vp]),
        "bb_backward_LMmu": (C.c_int, [vp, i32, i32, i32, vp, C.POINTER(Aux), vp, vp, vp, vp, vp]),
        "bb_guides_chain_nuH": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, C.POINTER(Aux), vp, vp, vp, dbl, pp, vp, vp,
                                          C.POINTER(dbl)]),
        "bb_guide_download_nuH": (C.c_int, [vp, vp, vp]),
        "bb_lptilde_nuH": (C.c_int, [vp, i32, vp, vp, dbl, vp, C.POINTER(dbl)]),
        "bb_lptilde_HV": (C.c_int, [vp, i32, i32, vp, vp, i32, vp, vp, vp, C.POINTER(dbl)]),
        "bb_guided_euler_ll": (C.c_int, [vp, C.POINTER(Model), pp, i32, u32]),
        "bb_llikelihood": (C.c_int, [vp, C.POINTER(Model), pp, i32]),
        "bb_innovations": (C.c_int, [vp, C.POINTER(Model), pp]),
        "bb_pcn_step": (C.c_int, [vp, C.POINTER(Model), pp, dbl, u64, u32, i32, u32]),
        "bb_ens_refresh_x": (C.c_int, [vp, C.POINTER(Model), pp]),
        "bb_ens_mc_reset": (C.c_int, [vp]),
        "bb_ens_mc_update": (C.c_int, [vp]),
        "bb_ens_mc_stats": (C.c_int, [vp, vp, vp, C.POINTER(i64)]),
        "bb_ens_chain_mc_reset": (C.c_int, [vp]),
        "bb_ens_chain_mc_update": (C.c_int, [vp]),
        "bb_ens_chain_mc_stats": (C.c_int, [vp, i64, i64, vp, vp, C.POINTER(i64)]),
        "bb_ens_chain_mc_band": (C.c_int, [vp, i64, i64, vp, vp]),
        "bb_pcn_step_host": (C.c_int, [vp, C.POINTER(Model), pp, dbl, u64, u32, i32, u32, vp, vp, vp, vp, vp]),
        "bb_theta_attach": (C.c_int, [vp, C.POINTER(Model), C.POINTER(ThetaSpec)]),
        "bb_theta_set": (C.c_int, [vp, i64, i64, vp]),
        "bb_theta_get": (C.c_int, [vp, C.c_int, i64, i64, vp]),
        "bb_theta_get_start": (C.c_int, [vp, C.c_int, i64, i64, vp]),
        "bb_theta_guides": (C.c_int, [vp]),
        "bb_theta_get_left": (C.c_int, [vp, C.c_int, i64, i64, vp]),
        "bb_theta_get_tables": (C.c_int, [vp, i64, vp, vp]),
        "bb_theta_guided_euler_ll": (C.c_int, [vp, i32, u32]),
        "bb_theta_pcn_step": (C.c_int, [vp, dbl, u64, u32, i32, u32]),
        "bb_theta_param_step": (C.c_int, [vp, vp, u64, u32, i32, u32]),
        "bb_theta_refresh_x": (C.c_int, [vp]),
        "bb_theta_block_step": (C.c_int, [vp, i32, i32, dbl, dbl, u64, u32, i32]),
        "bb_theta_get_block": (C.c_int, [vp, i64, i64, vp]),
        "bb_theta_get_acc": (C.c_int, [vp, C.POINTER(i64)]),
        "bb_theta_acc_device_ptr": (vp, [vp]),
        "bb_user_model_create": (C.c_int, [vp, i32, i32, C.c_char_p, C.POINTER(i32), C.POINTER(C.c_char_p), pp]),
        "bb_user_source_check": (C.c_int, [i32, i32, C.c_char_p, C.POINTER(i32), C.POINTER(C.c_char_p), i32, i32, i32, i32,
                                           C.c_char_p, i32]),
        "bb_user_model_handle": (i32, [vp]),
        "bb_user_model_log": (C.c_char_p, [vp]),
        "bb_user_model_destroy": (C.c_int, [vp]),
        "bb_comm_unique_id": (C.c_int, [vp]),
        "bb_comm_create": (C.c_int, [vp, i32, i32, vp, pp]),
        "bb_comm_adopt": (C.c_int, [vp, vp, i32, i32, pp]),
        "bb_comm_destroy": (C.c_int, [vp]),
        "bb_comm_rank": (C.c_int, [vp]),
        "bb_comm_size": (C.c_int, [vp]),
        "bb_comm_last_error": (C.c_char_p, []),
        "bb_allreduce_acc": (C.c_int, [vp, vp]),
        "bb_allreduce_theta_acc": (C.c_int, [vp, vp]),
        "bb_comm_get_acc": (C.c_int, [vp, C.POINTER(i64)]),
        "bb_comm_synchronize": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L, sorted(sig)


lib, SYMBOLS = _load()


def check(status: int) -> None:
    if status != 0:
        text = lib.bb_strerror(status).decode()
        if status == ERR_CUDA or status == ERR_NOMEM:
            text += ": " + lib.bb_last_cuda_error().decode()
        if status == ERR_COMM or status == ERR_UNSUPPORTED:
            extra = lib.bb_comm_last_error().decode()
            if extra:
                text += ": " + extra
        raise BridgeError(status, text)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
