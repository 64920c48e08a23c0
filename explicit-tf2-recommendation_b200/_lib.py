"""ctypes binding of libetr.so (the C ABI declared in include/etr.h).

There is NO fallback: if the shared library is missing, or a call fails, an
exception is raised.  The product path never touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libetr.so")

ETR_OK, ETR_EINVAL, ETR_ERANGE, ETR_ECUDA, ETR_ENOMEM, ETR_EUNSUPPORTED, ETR_EOVERFLOW, ETR_ETIMEOUT = range(8)
ETR_F32, ETR_BF16 = 0, 1
POOL_SUM, POOL_MEAN = 0, 1
ACT = {None: 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}
ADAM_ROWWISE, ADAM_KERAS_DENSE = 0, 1
ETR_TABLE_RECORD = -1          # etr_table.reserved: [var | m | v] interleaved in one 256-byte record


class EtrError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libetr status {status}: {msg}")
        self.status = status


class EtrIdRangeError(EtrError, IndexError):
    """An embedding id was out of range (TF-CPU raises InvalidArgumentError)."""


class EtrOverflowError(EtrError):
    """A sharded step dropped rows: a mailbox region or the owner's touched list overflowed."""


class EtrPeerTimeoutError(EtrError, TimeoutError):
    """A device-side peer barrier timed out."""


class etr_table(C.Structure):
    _fields_ = [("d_data", C.c_void_p), ("rows", C.c_int64), ("width", C.c_int32), ("stride", C.c_int32),
                ("dtype", C.c_int32), ("reserved", C.c_int32)]


class etr_ids(C.Structure):
    _fields_ = [("d_ids", C.c_void_p), ("d_csr_offsets", C.c_void_p), ("batch", C.c_int64),
                ("fields", C.c_int32), ("bag", C.c_int32),
                ("stride_b", C.c_int64), ("stride_f", C.c_int64), ("stride_l", C.c_int64),
                ("pad_id", C.c_int64), ("has_pad", C.c_int32), ("pooling", C.c_int32)]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_T, _I = C.POINTER(etr_table), C.POINTER(etr_ids)

# name -> (restype, argtypes); must list every function include/etr.h declares
PROTOTYPES = {
    "etr_version": (C.c_int, []),
    "etr_last_error": (C.c_char_p, []),
    "etr_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "etr_ctx_destroy": (C.c_int, [_vp]),
    "etr_ctx_poll_error": (C.c_int, [_vp, _vp, C.POINTER(_i64)]),
    "etr_ctx_launch_count": (_i64, [_vp]),
    "etr_ctx_peek_error_async": (C.c_int, [_vp, _vp, _vp]),
    "etr_ctx_decode_error": (C.c_int, [_vp, _vp, _vp, C.POINTER(_i64)]),
    "etr_assemble_ids": (C.c_int, [_vp, C.POINTER(_vp), _i32, _i64, _vp, _vp]),
    "etr_gather_fm_forward": (C.c_int, [_vp, _T, _i32, _i32, _I, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _i64, _i64, _vp]),
    "etr_embedding_gather": (C.c_int, [_vp, _T, _vp, _i64, _vp, _i64, _vp]),
    "etr_gather_fm_backward": (C.c_int, [_vp, _T, _i32, _i32, _I, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _vp]),
    "etr_sparse_plan_slots": (_i64, [_I, _i64]),
    "etr_sparse_plan": (C.c_int, [_vp, _I, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "etr_sparse_plan_keys": (C.c_int, [_vp, _I, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "etr_fm_fused_flat_apply": (C.c_int, [_vp, _T, _i32, _i32, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp,
                                          _f32, _f32, _f32, _vp]),
    "etr_sparse_segment_reduce": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "etr_sparse_segment_reduce_flat": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "etr_sparse_adam_apply": (C.c_int, [_vp, _T, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _f32, _vp, _f32, _f32, _f32, _i32, _vp]),
    "etr_fm_fused_backward_apply": (C.c_int, [_vp, _T, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                                              _i32, _i64, _i32, _f32, _vp, _f32, _f32, _f32, _i32, _vp, _vp]),
    "etr_fm_fused_prepare_bytes": (_i64, [_i64]),
    "etr_fm_fused_prepare": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "etr_fm_fused_backward_apply_prepared": (C.c_int, [_vp, _T, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                                                       _i32, _i64, _i32, _f32, _vp, _f32, _f32, _f32, _vp, _i64, _vp]),
    "etr_fm_fused_backward_push": (C.c_int, [_vp, _T, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _i32,
                                             _vp, _i32, C.POINTER(_vp), _vp, _i64, _vp]),
    "etr_dense_adam_apply": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _f32, _f32, _f32, _vp]),
    "etr_adam_step_begin": (C.c_int, [_vp, _vp, _f32, _f32, _f32, _vp]),
    "etr_bce_forward_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "etr_add_sigmoid": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "etr_gemm_f32": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _i64, _f32, _vp, _i64, _vp, _i64, _f32, _vp, _i64, _vp, _i32, _vp]),
    "etr_act_backward": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "etr_colsum_f32": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "etr_field_pair_forward": (C.c_int, [_vp, _T, _i32, _i32, _I, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "etr_field_pair_backward": (C.c_int, [_vp, _i32, _i32, _I, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "etr_pnn_forward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "etr_pnn_backward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "etr_pnn_kernel_grad": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _vp]),
    "etr_pair_dense_forward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "etr_pair_dense_backward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "etr_cross_vec_forward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "etr_cross_vec_backward": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "etr_cross_vec_finish": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "etr_cross_mat_layer_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "etr_cross_mat_layer_bf16": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp]),
    "etr_gemm_bf16_tn": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _vp]),
    "etr_gemm_bf16_tn_residual": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "etr_gemm_bf16_nn_wgrad": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "etr_gemm_bf16_tn_accumulate": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "etr_cross_mat_bwd_elementwise_bf16": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp]),
    "etr_cross_mat_bwd_du_colsum_bf16": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp]),
    "etr_cross_mat_bwd_dx0_bf16": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "etr_outer_bf16": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "etr_add_bf16_into_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "etr_colsum_bf16": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "etr_deepfm_tail_train": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    "etr_mlp_skinny_backward": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "etr_cast_bf16": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i32, _vp]),
    "etr_transpose_bf16": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "etr_peer_alloc": (C.c_int, [_vp, _i64, C.POINTER(_vp), _vp]),
    "etr_peer_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "etr_peer_close": (C.c_int, [_vp, _vp]),
    "etr_peer_free": (C.c_int, [_vp, _vp]),
    "etr_shard_set_create": (C.c_int, [_vp, C.POINTER(_vp), _i32, _i32, _i64, C.POINTER(_i32)]),
    "etr_shard_push": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, C.POINTER(_vp), C.POINTER(_vp),
                                 C.POINTER(_vp), _vp, _vp]),
    "etr_shard_request": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp]),
    "etr_shard_serve": (C.c_int, [_vp, _T, _vp, _vp, _i32, _i32, C.POINTER(_vp), _i32, _vp]),
    "etr_shard_vid_map": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "etr_shard_mailbox_accumulate": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "etr_shard_touched_adam": (C.c_int, [_vp, _T, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _f32, _f32, _f32, _vp]),
    "etr_shard_owner_prep": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "etr_shard_owner_apply": (C.c_int, [_vp, _T, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _f32, _f32,
                                        _f32, _vp]),
    "etr_peer_barrier": (C.c_int, [_vp, C.POINTER(_vp), _vp, _vp, _i32, _i32, _vp]),
    "etr_peer_allreduce_push": (C.c_int, [_vp, _vp, _i64, C.POINTER(_vp), _i32, _i32, _vp]),
    "etr_peer_allreduce_sum": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "etr_shard_partition": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "etr_cross_mat_bwd_elementwise": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "etr_used_rows_l2": (C.c_int, [_vp, _T, _vp, _vp, _i64, _f32, _vp, _i32, _vp, _vp]),
    "etr_sequence_pool_forward": (C.c_int, [_vp, _T, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "etr_sequence_pool_backward": (C.c_int, [_vp, _T, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "etr_tfrecord_parse": (C.c_int, [_vp, _i64, _i32, C.POINTER(C.c_char_p), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_vp),
                                     _i64, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen libetr.so and install the prototypes.  Raises if it is missing --
    build it with ``python -c 'import __graft_entry__ as g; g.build()'``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: the CUDA library is not built (run __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)         # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == ETR_OK:
        return
    msg = load().etr_last_error().decode("utf-8", "replace")
    if status == ETR_ERANGE:
        raise EtrIdRangeError(status, msg)
    if status == ETR_EOVERFLOW:
        raise EtrOverflowError(status, msg)
    if status == ETR_ETIMEOUT:
        raise EtrPeerTimeoutError(status, msg)
    raise EtrError(status, msg)
