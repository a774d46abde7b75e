"""ctypes binding of libs2vt_b200.so (the C ABI declared in include/s2vt_b200.h).

The product path has no fallback: if the shared library is missing or a call fails this module
raises.  Tensors are passed as raw device pointers; PyTorch only owns the memory and the stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2vt_b200.so")


class RowMap(C.Structure):
    _fields_ = [("inner", C.c_int32), ("stride_outer", C.c_int64), ("stride_inner", C.c_int64)]


def rowmap(inner: int, so: int, si: int) -> RowMap:
    return RowMap(int(inner), int(so), int(si))


def dense(ld: int) -> RowMap:
    return RowMap(1, int(ld), 0)


class XdecCfg(C.Structure):
    """s2vt_xdec_cfg (include/s2vt_b200.h)"""
    _fields_ = [(n, C.c_int32) for n in ("vocab_size", "feat_dim", "length", "dim_hid", "dim_embed", "sos_ix", "eos_ix")]


_vp, _i, _i64, _f, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32

# name -> (restype, argtypes); kept in one table so tests can compare it with the header
SIGNATURES = {
    "s2vt_abi_version": (_i, []),
    "s2vt_last_error": (C.c_char_p, []),
    "s2vt_launch_count": (_i64, []),
    "s2vt_add_launch_count": (_i64, [_i64]),
    "s2vt_adam_prepare": (_i, [_vp, _vp, _vp, _f, _f, _vp]),
    "s2vt_adam_f32_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _vp, _f, _vp]),
    "s2vt_has_tcgen05": (_i, []),
    "s2vt_device_error_flag": (_i, [_vp]),
    "s2vt_device_error_clear": (_i, []),
    "s2vt_gemm_f32": (_i, [_vp, _i, _i, _i, _vp, RowMap, _i, _vp, RowMap, _i, _vp, RowMap, _vp, _i, _i, _i64]),
    "s2vt_gemm_bf16": (_i, [_vp, _i, _i, _i, _vp, _i64, _i, _vp, _i64, _i, _vp, RowMap, _i, _vp, _i]),
    "s2vt_gemm_bf16_set_mode": (_i, [_i, _i]),
    "s2vt_cast_bf16": (_i, [_vp, _vp, _vp, _vp, _i64, _i64]),
    "s2vt_lstm_ws_bytes": (_i64, [_i, _i]),
    "s2vt_lstm_fwd_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_lstm_bf16_batch_pad": (_i64, [_i]),
    "s2vt_lstm_fwd_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_lstm_bwd_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_lstm_fwd_bf16_dir": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "s2vt_lstm_bwd_bf16_dir": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "s2vt_lstm_bf16_set_tiles_per_cluster": (_i, [_i]),
    "s2vt_lstm_fwd_bf16_sync": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _u32]),
    "s2vt_lstm_bwd_bf16_sync": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _u32]),
    "s2vt_set_launch_priority": (_i, [_i]),
    "s2vt_set_bulk_cta_cap": (_i, [_i]),
    "s2vt_gemm_bf16_gated": (_i, [_vp, _i, _i, _i, _vp, _i64, _vp, _i64, _i, _vp, _i64, _vp, _i, _i, _i, _i, _vp, _vp, _u32, _vp, _vp]),
    "s2vt_graph_instantiate": (_i, [_vp, _i, _vp]),
    "s2vt_graph_launch": (_i, [_vp, _vp]),
    "s2vt_graph_exec_destroy": (_i, [_vp]),
    "s2vt_graph_kernel_priorities": (_i, [_vp, _vp, _i, _vp]),
    "s2vt_timestamp": (_i, [_vp, _vp]),
    "s2vt_stream_wait_value32": (_i, [_vp, _vp, _u32]),
    "s2vt_stream_write_value32": (_i, [_vp, _vp, _u32]),
    "s2vt_lstm_bwd_bf16_chunk": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i]),
    "s2vt_lstm_bwd_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_embed_gather_f32": (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _vp, _i64]),
    "s2vt_embed_scatter_add_f32": (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _vp, _i64]),
    "s2vt_embed_gather_bf16": (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _vp, _i64]),
    "s2vt_colsum_bf16": (_i, [_vp, _vp, _i64, _i, _i64, _vp, _vp]),
    "s2vt_ce_bf16": (_i, [_vp, _vp, _i64, _i, _vp, RowMap, _vp, _vp, _vp, _i, _vp, _vp]),
    "s2vt_ce_bf16_mapped": (_i, [_vp, _vp, _i64, _i, _vp, RowMap, _vp, _vp, _vp, _i, _vp, RowMap, _vp]),
    "s2vt_bcast_rows_f32": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "s2vt_vocab_ce_ws_bytes": (_i64, [_i, _i]),
    "s2vt_vocab_ce_fwd_bf16": (_i, [_vp, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, RowMap, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_ce_dlogits_inplace_bf16": (_i, [_vp, _vp, _i64, _i, _i64, _vp, _vp, RowMap, _vp]),
    "s2vt_add_f32": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "s2vt_colsum_f32": (_i, [_vp, _vp, _i64, _i, _i64, _vp, _i]),
    "s2vt_ce_f32": (_i, [_vp, _vp, _i64, _i, _vp, RowMap, _vp, _vp, _vp, _vp]),
    "s2vt_adam_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _f, _vp]),
    "s2vt_greedy_ws_bytes": (_i64, [_i, _i, _i, _i]),
    "s2vt_greedy_decode_f32": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_beam_ws_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "s2vt_beam_search_f32": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp]),
    "s2vt_lstm_steps_fwd_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_lstm_steps_bwd_ws_bytes": (_i64, [_i, _i]),
    "s2vt_lstm_steps_bwd_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2vt_xgemm_ws_bytes": (_i64, [_i, _i, _i]),
    "s2vt_xgemm_f32": (_i, [_vp, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, RowMap, _vp, _i, _vp]),
    "s2vt_xdec_set_trace": (_i, [_vp, _i]),
    "s2vt_xdec_weights_bytes": (_i64, [XdecCfg]),
    "s2vt_xdec_prepare": (_i, [_vp, XdecCfg, _vp, _vp]),
    "s2vt_xdec_greedy_ws_bytes": (_i64, [XdecCfg, _i]),
    "s2vt_xdec_greedy": (_i, [_vp, XdecCfg, _vp, _i, _vp, _vp, _vp]),
    "s2vt_xdec_beam_ws_bytes": (_i64, [XdecCfg, _i, _i, _i]),
    "s2vt_xdec_beam": (_i, [_vp, XdecCfg, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
}

_lib: Optional[C.CDLL] = None


class S2VTLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises S2VTLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S2VTLibraryError(
            "libs2vt_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python s2vt-video-caption_b200/build.py`; there is no CPU / PyTorch fallback on the product path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise S2VTLibraryError("libs2vt_b200.so does not export %s (stale build?)" % name)
        fn.restype = res
        fn.argtypes = args
    if lib.s2vt_abi_version() != 1:
        raise S2VTLibraryError("ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().s2vt_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise S2VTLibraryError("%s failed: %s" % (what, last_error()))


def launch_count() -> int:
    return int(load().s2vt_launch_count())


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor], offset: int = 0) -> Optional[int]:
    """Raw device address of element `offset` of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr() + offset * t.element_size()


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "s2vt_b200 runs on CUDA (sm_100a) only: got a %s tensor. There is no CPU fallback; "
                "move the model and inputs to a B200 (`.cuda()`)." % t.device)
