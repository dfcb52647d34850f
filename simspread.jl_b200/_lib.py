"""ctypes binding of libsimspread_b200.so -- argument lists identical to include/simspread_b200.h
(and to the `ccall`s of simspread.jl_b200/julia/SimSpreadB200.jl).  There is no CPU fallback: if
the library cannot be loaded, or no B200 is visible, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import re

from . import _build

c_i32, c_i64, c_u32, c_f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_double
vp = C.c_void_p
P = C.POINTER

SS_OK, SS_ERR_INVALID, SS_ERR_CUDA, SS_ERR_NO_DEVICE, SS_ERR_ASSERT, SS_ERR_OOM, SS_ERR_UNSUPPORTED = range(7)
SS_PREDICT_CLEAN = 1
SS_OP_N, SS_OP_T = 0, 1
SS_PRECISION_F64, SS_PRECISION_TF32, SS_PRECISION_F64_INT8 = 0, 1 << 4, 3 << 4


class SimSpreadError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libsimspread_b200 status {status}: {msg}")
        self.status = status
        self.message = msg


# name -> (restype, argtypes); kept in the order of the header
SIGNATURES = {
    "ss_version": (c_i32, []),
    "ss_last_error": (C.c_char_p, []),
    "ss_device_count": (c_i32, [P(c_i32)]),
    "ss_ctx_create": (c_i32, [c_i32, P(vp)]),
    "ss_ctx_destroy": (c_i32, [vp]),
    "ss_ctx_sync": (c_i32, [vp]),
    "ss_ctx_stream": (c_i32, [vp, P(vp)]),
    "ss_ctx_launch_count": (c_i32, [vp, P(c_i64)]),
    "ss_ctx_int8_stats": (c_i32, [vp, vp]),
    "ss_ctx_profile": (c_i32, [vp, c_i32]),
    "ss_ctx_profile_read": (c_i32, [vp, P(c_f64), P(c_f64), c_i32, P(c_i32)]),
    "ss_host_alloc": (c_i32, [c_i64, P(vp)]),
    "ss_host_free": (c_i32, [vp]),
    "ss_mat_create": (c_i32, [vp, c_i64, c_i64, P(vp)]),
    "ss_mat_create_ipc": (c_i32, [vp, c_i64, c_i64, P(vp)]),
    "ss_mat_wrap": (c_i32, [vp, vp, c_i64, c_i64, c_i64, P(vp)]),
    "ss_mat_destroy": (c_i32, [vp]),
    "ss_mat_info": (c_i32, [vp, P(c_i64), P(c_i64), P(c_i64), P(vp)]),
    "ss_mat_upload": (c_i32, [vp, vp, vp, c_i64]),
    "ss_mat_upload_rowmajor": (c_i32, [vp, vp, vp, c_i64]),
    "ss_mat_download": (c_i32, [vp, vp, vp, c_i64]),
    "ss_mat_upload_cols_async": (c_i32, [vp, vp, c_i64, c_i64, vp, c_i64]),
    "ss_mat_download_cols_async": (c_i32, [vp, vp, c_i64, c_i64, vp, c_i64]),
    "ss_ivec_create": (c_i32, [vp, c_i64, P(vp)]),
    "ss_ivec_wrap": (c_i32, [vp, vp, c_i64, P(vp)]),
    "ss_ivec_destroy": (c_i32, [vp]),
    "ss_ivec_info": (c_i32, [vp, P(c_i64), P(vp)]),
    "ss_ivec_upload": (c_i32, [vp, vp, vp]),
    "ss_ivec_download": (c_i32, [vp, vp, vp]),
    "ss_text_matrix_dims": (c_i32, [C.c_char_p, c_i32, P(c_i64), P(c_i64)]),
    "ss_text_matrix_read": (c_i32, [C.c_char_p, c_i32, c_i32, c_i32, vp, c_i64, c_i64, c_i64]),
    "ss_save_rows": (c_i32, [C.c_char_p, c_i32, c_i64, c_i64, c_i64, vp, vp, vp, c_i64, c_i32, vp, c_i64, c_i32, c_i32, P(c_i64)]),
    "ss_save_rows_mat": (c_i32, [vp, C.c_char_p, c_i32, c_i64, vp, vp, vp, vp, c_i32, c_i32, P(c_i64)]),
    "ss_featurize": (c_i32, [vp, vp, c_f64, c_i32, vp]),
    "ss_featurize_csr": (c_i32, [vp, vp, c_f64, c_i32, P(vp)]),
    "ss_featurize_csc": (c_i32, [vp, vp, c_f64, c_i32, P(vp)]),
    "ss_jaccard_featurize": (c_i32, [vp, vp, vp, c_f64, c_i32, vp]),
    "ss_tanimoto_featurize_bits": (c_i32, [vp, vp, c_i64, vp, c_i64, c_i64, c_f64, c_i32, vp]),
    "ss_csr_info": (c_i32, [vp, P(c_i64), P(c_i64), P(c_i64), P(c_i32)]),
    "ss_csr_download": (c_i32, [vp, vp, vp, vp, vp]),
    "ss_csr_destroy": (c_i32, [vp]),
    "ss_csr_wrap": (c_i32, [vp, c_i64, c_i64, c_i64, vp, vp, vp, P(vp)]),
    "ss_recommend_topl": (c_i32, [vp, vp, vp, c_i32, c_i64, c_i64, vp, vp]),
    "ss_comm_unique_id": (c_i32, [vp]),
    "ss_comm_init": (c_i32, [vp, c_i32, c_i32, vp, P(vp)]),
    "ss_comm_init_file": (c_i32, [vp, c_i32, c_i32, C.c_char_p, c_f64, P(vp)]),
    "ss_comm_destroy": (c_i32, [vp]),
    "ss_comm_info": (c_i32, [vp, P(c_i32), P(c_i32), P(c_i32)]),
    "ss_comm_barrier": (c_i32, [vp]),
    "ss_comm_allreduce_i32": (c_i32, [vp, vp]),
    "ss_comm_allgather_i32": (c_i32, [vp, vp, vp]),
    "ss_comm_allgather_host": (c_i32, [vp, vp, vp, c_i64]),
    "ss_comm_allgather_cols": (c_i32, [vp, vp]),
    "ss_comm_allreduce_host_f64": (c_i32, [vp, vp, c_i32, c_i32]),
    "ss_sharded_create": (c_i32, [vp, c_i64, c_i64, c_i64, P(vp)]),
    "ss_sharded_destroy": (c_i32, [vp]),
    "ss_sharded_info": (c_i32, [vp, P(c_i64), P(c_i32)]),
    "ss_sharded_front": (c_i32, [vp, vp, vp]),
    "ss_sharded_views": (c_i32, [vp, P(vp), P(vp)]),
    "ss_predict_query_sharded": (c_i32, [vp, vp, vp, vp, vp, c_u32]),
    "ss_transfer_build": (c_i32, [vp, vp, vp, P(vp)]),
    "ss_transfer_info": (c_i32, [vp, vp]),
    "ss_transfer_download": (c_i32, [vp, vp, vp, vp]),
    "ss_transfer_destroy": (c_i32, [vp]),
    "ss_recommend_topl_transfer": (c_i32, [vp, vp, vp, c_i32, c_i64, c_i64, vp, vp]),
    "ss_gather": (c_i32, [vp, vp, vp, vp, vp]),
    "ss_degrees": (c_i32, [vp, vp, vp, vp, vp, vp]),
    "ss_k_rows": (c_i32, [vp, vp, vp]),
    "ss_spread_rows": (c_i32, [vp, vp, vp, vp]),
    "ss_gemm_f64": (c_i32, [vp, c_i32, vp, vp, vp, vp, vp]),
    "ss_gemm_lowp": (c_i32, [vp, c_i32, vp, vp, vp, vp, vp, c_u32]),
    "ss_gemm_f64_mirrored": (c_i32, [vp, c_i32, vp, vp, vp, vp, vp, c_i32, P(vp)]),
    "ss_mat_ipc_handle": (c_i32, [vp, vp, vp]),
    "ss_ipc_open": (c_i32, [vp, vp, P(vp)]),
    "ss_ipc_close": (c_i32, [vp, vp]),
    "ss_predict_query": (c_i32, [vp, vp, vp, vp, vp, c_u32, vp]),
    "ss_predict_query_fetch": (c_i32, [vp, vp, vp, vp, vp, c_u32, vp, c_i64]),
    "ss_predict_query_folds": (c_i32, [vp, vp, vp, c_i32, vp, vp, vp, vp, vp, vp, vp, vp, c_u32]),
    "ss_predict_query_csr": (c_i32, [vp, vp, vp, vp, vp, c_u32, vp]),
    "ss_predict_source": (c_i32, [vp, vp, vp, vp, c_u32]),
    "ss_clean": (c_i32, [vp, vp, vp]),
    "ss_predict_query_host": (c_i32, [vp, vp, c_i64, vp, c_i64, vp, c_i64, c_i64, c_i64, c_i64, c_i64,
                                      c_u32, vp, c_i64]),
    "ss_stream_product_host": (c_i32, [vp, vp, c_i64, c_i64, vp, vp, vp, c_i64]),
    "ss_topl_rows": (c_i32, [vp, vp, c_i32, vp, vp]),
    "ss_atl": (c_i32, [vp, vp, vp, c_i32, P(c_f64)]),
    "ss_auroc_auprc": (c_i32, [vp, vp, vp, c_i64, P(c_f64)]),
    "ss_auc_sort": (c_i32, [vp, vp, vp, vp, c_i64, P(vp), P(vp)]),
    "ss_auc_lower_bound": (c_i32, [vp, vp, c_i64, vp, c_i32, vp]),
    "ss_auc_segment_summary": (c_i32, [vp, vp, vp, c_i64, vp]),
    "ss_auc_segment_integrate": (c_i32, [vp, vp, vp, c_i64, vp, vp]),
    "ss_auroc_auprc_mat": (c_i32, [vp, vp, vp, P(c_f64)]),
    "ss_bedroc": (c_i32, [vp, vp, vp, c_i32, c_f64, P(c_f64)]),
    "ss_threshold_sweep": (c_i32, [vp, vp, vp, c_i32, P(c_f64)]),
}

_lib = None


def header_symbols():
    """Every function the public header declares (used by the CPU-side export test)."""
    hdr = os.path.join(os.path.dirname(_build.HERE), "include", "simspread_b200.h")
    with open(hdr) as f:
        return re.findall(r"^SS_API\s+[\w\s\*]+?\b(ss_[a-z0-9_]+)\(", f.read(), flags=re.M)


def lib():
    """Load (once) libsimspread_b200.so.  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        path = _build.lib_path()
        if not os.path.exists(path):
            raise SimSpreadError(-1, f"{path} is missing: run `python __graft_entry__.py` (build()) "
                                     "first; simspread_b200 has no CPU fallback")
        h = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status: int):
    if status == SS_OK:
        return
    msg = lib().ss_last_error().decode("utf-8", "replace")
    if status == SS_ERR_ASSERT:
        raise AssertionError(msg)  # mirrors Julia's AssertionError(msg) of the reference @assert
    raise SimSpreadError(status, msg)
