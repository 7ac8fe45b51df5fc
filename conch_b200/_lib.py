"""ctypes binding of libconch_b200.so -- the only way the Python side reaches the GPU.

Mirrors include/conch_b200.h one to one.  Loading fails loudly: there is no fallback implementation.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import threading

_PKG = pathlib.Path(__file__).resolve().parent
# CB200_DEBUG_LIB=1 selects the instrumented build made with CB200_NVCC_EXTRA (benchmarks/tc_trace.py); still our CUDA library
LIB_PATH = _PKG / ("libconch_b200_dbg.so" if os.environ.get("CB200_DEBUG_LIB") == "1" else "libconch_b200.so")

# enums of include/conch_b200.h
F16, BF16, F32 = 0, 1, 2
ZP_NONE, ZP_I32, ZP_U8, ZP_ACT = 0, 1, 2, 3
IN_INT8, IN_FP8_E4M3 = 0, 1
AUTO, DECODE_MMA, TC_SS, TC_TS, TC2_TS, DECODE_STREAM = 0, 1, 2, 3, 4, 5
VARIANT_NAMES = {AUTO: "auto", DECODE_MMA: "decode_mma", TC_SS: "tc_ss", TC_TS: "tc_ts", TC2_TS: "tc2_ts", DECODE_STREAM: "decode_stream"}
ERR_ARG, ERR_SHAPE, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_DEVICE = -1, -2, -3, -4, -5
WORKSPACE_COUNTER_BYTES = 65536

# every symbol include/conch_b200.h declares (tests check the .so exports exactly these)
EXPORTS = (
    "cb200_version",
    "cb200_last_error",
    "cb200_mpgemm_workspace_bytes",
    "cb200_mpgemm",
    "cb200_mpgemm_bias",
    "cb200_scaled_gemm",
    "cb200_static_scaled_quant",
    "cb200_mpgemm_push",
    "cb200_mpgemm_push_signal",
    "cb200_wait_flags",
    "cb200_mpgemm_chain",
    "cb200_push",
    "cb200_push_bulk",
    "cb200_peer_copy",
    "cb200_last_variant",
    "cb200_last_launch_count",
    "cb200_dequantize",
    "cb200_set_decode_tuning",
    "cb200_set_tc_tuning",
    "cb200_set_pdl",
    "cb200_set_tc_cluster",
    "cb200_set_stream_tuning",
    "cb200_debug_set_stream_trace",
)

_lock = threading.Lock()
_lib: ctypes.CDLL | None = None


class LibraryNotBuiltError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the C ABI.  Raises if the library has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise LibraryNotBuiltError(
                f"{LIB_PATH} is missing. Build it with `python -m conch_b200.build` (needs nvcc). "
                "conch_b200 has no CPU or Triton fallback."
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        c_void_p, c_int, c_int64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
        lib.cb200_version.restype = c_int
        lib.cb200_version.argtypes = []
        lib.cb200_last_error.restype = ctypes.c_char_p
        lib.cb200_last_error.argtypes = []
        lib.cb200_last_variant.restype = c_int
        lib.cb200_last_variant.argtypes = []
        lib.cb200_last_launch_count.restype = c_int
        lib.cb200_last_launch_count.argtypes = []
        lib.cb200_mpgemm_workspace_bytes.restype = c_size_t
        lib.cb200_mpgemm_workspace_bytes.argtypes = [c_int64, c_int64, c_int64, c_int, c_int]
        lib.cb200_mpgemm.restype = c_int
        lib.cb200_mpgemm.argtypes = [
            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,  # x, w_q_packed, w_s, w_zp, out
            c_int64, c_int64, c_int64,  # M, K, N
            c_int, c_int, c_int64,  # bits, bias, group_size
            c_int, c_int, c_int, c_int,  # act_dtype, out_dtype, zp_kind, zp_is_scalar
            c_void_p, c_size_t,  # workspace
            c_int, c_void_p,  # variant, stream
        ]  # fmt: skip
        lib.cb200_mpgemm_bias.restype = c_int
        lib.cb200_mpgemm_bias.argtypes = lib.cb200_mpgemm.argtypes[:4] + [c_void_p] + lib.cb200_mpgemm.argtypes[4:]  # ..., w_zp, bias, out, ...
        lib.cb200_scaled_gemm.restype = c_int
        lib.cb200_scaled_gemm.argtypes = [c_void_p] * 6 + [c_int64] * 3 + [c_int] * 4 + [c_void_p]  # a, b_t, sa, sb, bias, out, M, K, N, ...
        lib.cb200_static_scaled_quant.restype = c_int
        lib.cb200_static_scaled_quant.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]
        lib.cb200_mpgemm_push.restype = c_int
        lib.cb200_mpgemm_push.argtypes = lib.cb200_mpgemm.argtypes + [ctypes.POINTER(c_void_p), c_int]  # peer_out, num_peers
        lib.cb200_mpgemm_push_signal.restype = c_int
        lib.cb200_mpgemm_push_signal.argtypes = lib.cb200_mpgemm_push.argtypes + [ctypes.POINTER(c_void_p), c_void_p]
        lib.cb200_mpgemm_chain.restype = c_int
        lib.cb200_mpgemm_chain.argtypes = lib.cb200_mpgemm_push_signal.argtypes + [c_void_p, c_int, c_void_p]  # x_flags, count, x_expected
        lib.cb200_push.restype = c_int
        lib.cb200_push.argtypes = [c_void_p, c_size_t, ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_void_p), c_void_p, c_void_p]
        lib.cb200_push_bulk.restype = c_int
        lib.cb200_push_bulk.argtypes = [c_void_p, c_size_t, ctypes.POINTER(c_void_p), c_int, c_void_p, c_int, c_void_p]
        lib.cb200_peer_copy.restype = c_int
        lib.cb200_peer_copy.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        lib.cb200_wait_flags.restype = c_int
        lib.cb200_wait_flags.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
        lib.cb200_dequantize.restype = c_int
        lib.cb200_dequantize.argtypes = [
            c_void_p, c_void_p, c_void_p, c_void_p,
            c_int64, c_int64,
            c_int, c_int, c_int64,
            c_int, c_int, c_int, c_void_p,
        ]  # fmt: skip
        lib.cb200_set_decode_tuning.restype = None
        lib.cb200_set_decode_tuning.argtypes = [c_int, c_int, c_int, c_int]
        lib.cb200_set_tc_tuning.restype = None
        lib.cb200_set_tc_tuning.argtypes = [c_int, c_int]
        lib.cb200_set_tc_cluster.restype = None
        lib.cb200_set_tc_cluster.argtypes = [c_int]
        lib.cb200_set_stream_tuning.restype = None
        lib.cb200_set_stream_tuning.argtypes = [c_int, c_int]
        lib.cb200_debug_set_stream_trace.restype = None
        lib.cb200_debug_set_stream_trace.argtypes = [c_void_p]
        lib.cb200_set_pdl.restype = None
        lib.cb200_set_pdl.argtypes = [c_int]
        _lib = lib
    return _lib


def raise_for_status(rc: int) -> None:
    """0 -> ok; <0 -> ValueError / NotImplementedError (reference's exception types); >0 -> CUDA RuntimeError."""
    if rc == 0:
        return
    msg = (load().cb200_last_error() or b"").decode(errors="replace")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc in (ERR_ARG, ERR_SHAPE, ERR_WORKSPACE):
        raise ValueError(msg)
    raise RuntimeError(f"libconch_b200 error {rc}: {msg}")
