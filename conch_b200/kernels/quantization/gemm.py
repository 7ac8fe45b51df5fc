"""Launcher layer: the B200 counterpart of the reference's `conch/kernels/quantization/gemm.py`.

The reference launcher (`mixed_precision_gemm_launcher`, conch/kernels/quantization/gemm.py:482-545) computes a
Triton grid and JIT-launches `_gemm_kernel`.  Here the launcher marshals raw device pointers into the C ABI
(`cb200_mpgemm`, include/conch_b200.h) which picks the decode or tcgen05 kernel.  Like the reference launcher
it does not allocate the output; the only memory it owns is a cached split-K workspace per (device, stream).
"""

from __future__ import annotations

import ctypes
from collections.abc import Sequence
from dataclasses import dataclass
from enum import Enum

import torch

from conch_b200 import _lib


class WeightGroupMode(Enum):
    """Dequantisation modes (names follow the reference enum, conch/kernels/quantization/gemm.py:34-41).

    Only the modes reachable from `mixed_precision_gemm` (conch/ops/quantization/gemm.py:34-38) and `scaled_gemm` (:227)
    exist here.
    """

    NONE = 0  # scaled_gemm: nothing to dequantise
    SYMMETRIC_NO_SHIFT = 2  # W = q * s            (w_zp is None)
    SYMMETRIC_WITH_SHIFT = 3  # W = (q - zp) * s


class ChannelScaleMode(Enum):
    """conch/kernels/quantization/gemm.py:68-74; the mixed-precision frontend always uses NONE, scaled_gemm always
    WEIGHT_AND_ACTIVATION."""

    NONE = 0
    WEIGHT_AND_ACTIVATION = 3


@dataclass
class MixedPrecisionMatmulMetadata:
    """Same 16 fields as the reference dataclass (conch/kernels/quantization/gemm.py:460-479)."""

    m_dim: int
    k_dim: int
    n_dim: int
    weight_size_bits: int
    weight_bias: int
    group_size: int
    elements_per_sample: int
    zero_is_scalar: bool
    unpack_mask: int
    data_contiguous: bool
    input_dtype: torch.dtype
    output_dtype: torch.dtype
    acc_dtype: torch.dtype
    meta_dtype: torch.dtype
    channel_scale_mode: ChannelScaleMode
    weight_group_mode: WeightGroupMode


@dataclass
class ScaledMatmulMetadata:
    """Same 10 fields as the reference dataclass (conch/kernels/quantization/gemm.py:548-561)."""

    m_dim: int
    k_dim: int
    n_dim: int
    data_contiguous: bool
    input_dtype: torch.dtype
    output_dtype: torch.dtype
    acc_dtype: torch.dtype
    meta_dtype: torch.dtype
    channel_scale_mode: ChannelScaleMode
    weight_group_mode: WeightGroupMode


_DTYPE_CODE = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}
_IN_KIND = {torch.int8: _lib.IN_INT8, torch.float8_e4m3fn: _lib.IN_FP8_E4M3}

# (device index, stream handle) -> uint8 workspace tensor whose counter prefix is zero between launches
_workspaces: dict[tuple[int, int], torch.Tensor] = {}
_need_cache: dict[tuple[int, int, int, int, int], int] = {}


def _raw_stream(index: int) -> int:
    """cudaStream_t of torch's current stream on device `index` (the private fast accessor when it exists)."""
    try:
        return torch._C._cuda_getCurrentRawStream(index)
    except AttributeError:  # pragma: no cover - older / newer torch without the private hook
        return torch.cuda.current_stream(index).cuda_stream


def _workspace(device: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        # must start zeroed: the first WORKSPACE_COUNTER_BYTES bytes are arrival counters that kernels
        # reset themselves, so this memset happens once per growth, never per call
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def zero_point_kind(x: torch.Tensor, w_zp: torch.Tensor | None) -> int:
    if w_zp is None:
        return _lib.ZP_NONE
    if w_zp.dtype == torch.int32:
        return _lib.ZP_I32
    if w_zp.dtype == torch.uint8:
        return _lib.ZP_U8
    if w_zp.dtype == x.dtype:
        return _lib.ZP_ACT
    raise NotImplementedError(f"w_zp dtype {w_zp.dtype} not supported (int32, uint8 or the dtype of x)")


def mixed_precision_gemm_launcher(
    output: torch.Tensor,
    x: torch.Tensor,
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    metadata: MixedPrecisionMatmulMetadata,
    *,
    variant: int = _lib.AUTO,
    peer_out_ptrs: Sequence[int] | None = None,
    signal: tuple[Sequence[int], int] | None = None,
    x_ready: tuple[int, int, int] | None = None,
    bias: torch.Tensor | None = None,
) -> None:
    """Enqueue out = x @ dequant(W) (+ bias) on the current CUDA stream.  Never synchronises.

    Counterpart of the reference launcher (conch/kernels/quantization/gemm.py:482-545): same argument order.
    `peer_out_ptrs` (multi-GPU, decode regime): device pointers of other GPUs' [M, N] buffers the kernel's epilogue also
    stores the result to (`cb200_mpgemm_push`, see conch_b200/sharded.py).  `signal` = (peer flag pointers, local flag pointer):
    completion counters raised once everything is pushed (`cb200_mpgemm_push_signal`).  `x_ready` = (flags pointer, count,
    expected pointer): `x` is a gather buffer that is complete once those counters have been raised -- the kernel itself waits
    (`cb200_mpgemm_chain`).
    """
    lib = _lib.load()
    device = x.device
    index = device.index if device.index is not None else torch.cuda.current_device()
    stream = _raw_stream(index)
    m, k, n = metadata.m_dim, metadata.k_dim, metadata.n_dim
    # the scratch size of a shape is cached (one ctypes call less per launch); the tuning knobs can change it, in which
    # case the library answers CB200_ERR_WORKSPACE and the size is queried again below
    key = (m, k, n, metadata.weight_size_bits, variant)
    need = _need_cache.get(key)
    if need is None:
        need = _need_cache[key] = lib.cb200_mpgemm_workspace_bytes(m, k, n, metadata.weight_size_bits, variant)
    ws = _workspace(device, stream, need)
    zp_kind = zero_point_kind(x, w_zp)
    args = (
        x.data_ptr(),
        w_q_packed.data_ptr(),
        w_s.data_ptr(),
        w_zp.data_ptr() if w_zp is not None else None,
        output.data_ptr(),
        m, k, n,
        metadata.weight_size_bits,
        metadata.weight_bias,
        metadata.group_size,
        _DTYPE_CODE[metadata.input_dtype],
        _DTYPE_CODE[metadata.output_dtype],
        zp_kind,
        int(metadata.zero_is_scalar),
        ws.data_ptr(),
        ws.numel(),
        variant,
        stream,
    )  # fmt: skip
    entry = lib.cb200_mpgemm
    ws_at = 15  # position of the workspace pointer in `args`
    if bias is not None:
        if peer_out_ptrs or signal is not None or x_ready is not None:
            raise NotImplementedError("bias and the multi-GPU push epilogue cannot be combined")
        entry = lib.cb200_mpgemm_bias  # (..., w_zp, bias, out, ...)
        args = args[:4] + (bias.data_ptr(),) + args[4:]
        ws_at = 16
    if peer_out_ptrs or signal is not None or x_ready is not None:
        peers = list(peer_out_ptrs or ())
        entry = lib.cb200_mpgemm_push
        args = args + ((ctypes.c_void_p * max(len(peers), 1))(*peers), len(peers))
        if signal is not None or x_ready is not None:
            # (peer flag pointers, local flag pointer): the kernel bumps the counters once the whole result is pushed
            peer_flags, local_flag = signal if signal is not None else ((), None)
            entry = lib.cb200_mpgemm_push_signal
            args = args + ((ctypes.c_void_p * max(len(peer_flags), 1))(*peer_flags), local_flag)
        if x_ready is not None:
            entry = lib.cb200_mpgemm_chain
            args = args + tuple(x_ready)
    if index != torch.cuda.current_device():
        with torch.cuda.device(device):
            rc = entry(*args)
    else:
        rc = entry(*args)
    if rc == _lib.ERR_WORKSPACE:  # stale cached size (a tuning knob changed the plan): ask again, grow, retry once
        need = _need_cache[key] = lib.cb200_mpgemm_workspace_bytes(m, k, n, metadata.weight_size_bits, variant)
        ws = _workspace(device, stream, need)
        args = args[:ws_at] + (ws.data_ptr(), ws.numel()) + args[ws_at + 2 :]
        with torch.cuda.device(device):
            rc = entry(*args)
    _lib.raise_for_status(rc)


def dequantize_launcher(
    output: torch.Tensor,
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    weight_size_bits: int,
    weight_bias: int,
    group_size: int,
) -> None:
    """Enqueue W = dequant(w_q_packed, w_s, w_zp): the reference's `_dequantize` stage on its own
    (conch/kernels/quantization/gemm.py:176-216)."""
    lib = _lib.load()
    device = w_q_packed.device
    stream = torch.cuda.current_stream(device).cuda_stream
    k, n = output.shape
    zp_kind = zero_point_kind(output, w_zp)
    with torch.cuda.device(device):
        rc = lib.cb200_dequantize(
            w_q_packed.data_ptr(),
            w_s.data_ptr(),
            w_zp.data_ptr() if w_zp is not None else None,
            output.data_ptr(),
            k, n,
            weight_size_bits, weight_bias, group_size,
            _DTYPE_CODE[output.dtype],
            zp_kind,
            int(w_zp is not None and w_zp.numel() == 1),
            stream,
        )  # fmt: skip
    _lib.raise_for_status(rc)


def _k_major(b: torch.Tensor) -> torch.Tensor:
    """`b` [K, N] as a dense [N, K] tensor (K contiguous): a free view when b is the transpose of a contiguous [N, K] weight --
    the layout the reference's callers use (tests/scaled_gemm_test.py:68-72) -- else one transposing copy."""
    bt = b.t()
    if not bt.is_contiguous() or bt.data_ptr() % 16:
        bt = bt.contiguous().clone() if bt.is_contiguous() else bt.contiguous()
    return bt


def scaled_gemm_launcher(
    output: torch.Tensor,
    a: torch.Tensor,
    b: torch.Tensor,
    scale_a: torch.Tensor,
    scale_b: torch.Tensor,
    metadata: ScaledMatmulMetadata,
    *,
    bias: torch.Tensor | None = None,
) -> None:
    """Enqueue out = cast((a @ b) * (scale_a * scale_b^T)) (+ bias) on the current CUDA stream.

    Counterpart of the reference launcher (conch/kernels/quantization/gemm.py:564-627): same argument order; the bias the
    reference op adds with a second kernel (conch/ops/quantization/gemm.py:249-250) can be handed to the epilogue instead,
    which reproduces that two-rounding arithmetic."""
    lib = _lib.load()
    if metadata.input_dtype not in _IN_KIND:
        raise NotImplementedError(f"scaled_gemm input dtype {metadata.input_dtype} (int8 or float8_e4m3fn)")
    device = a.device
    index = device.index if device.index is not None else torch.cuda.current_device()
    a = a.contiguous()
    if a.data_ptr() % 16:
        a = a.clone()
    bt = _k_major(b)
    sa = scale_a.to(torch.float32).contiguous().reshape(-1)
    sb = scale_b.to(torch.float32).contiguous().reshape(-1)
    with torch.cuda.device(device):
        rc = lib.cb200_scaled_gemm(
            a.data_ptr(), bt.data_ptr(), sa.data_ptr(), sb.data_ptr(), None if bias is None else bias.data_ptr(), output.data_ptr(),
            metadata.m_dim, metadata.k_dim, metadata.n_dim, _IN_KIND[metadata.input_dtype], _DTYPE_CODE[metadata.output_dtype],
            int(sa.numel() == 1), int(sb.numel() == 1), _raw_stream(index),
        )  # fmt: skip
    _lib.raise_for_status(rc)


def static_scaled_quant_launcher(output_tensor: torch.Tensor, input_tensor: torch.Tensor, scale: torch.Tensor) -> None:
    """Enqueue output = quantise(input / scale) for a whole tensor: int8 or float8_e4m3fn, by the dtype of `output_tensor`.
    Counterpart of static_scaled_int8_quant_launcher / static_scaled_fp8_quant_launcher
    (conch/kernels/quantization/int8.py:61-97, fp8.py:67-97)."""
    lib = _lib.load()
    if output_tensor.dtype not in _IN_KIND:
        raise NotImplementedError(f"quantised dtype {output_tensor.dtype} (int8 or float8_e4m3fn)")
    if input_tensor.dtype not in _DTYPE_CODE:
        raise NotImplementedError(f"input dtype {input_tensor.dtype} (fp16, bf16 or fp32)")
    if not output_tensor.is_contiguous():
        raise ValueError("output_tensor must be contiguous")
    device = input_tensor.device
    index = device.index if device.index is not None else torch.cuda.current_device()
    x = input_tensor.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    s = scale.to(torch.float32).contiguous()
    with torch.cuda.device(device):
        rc = lib.cb200_static_scaled_quant(
            x.data_ptr(), s.data_ptr(), output_tensor.data_ptr(), x.numel(), _DTYPE_CODE[x.dtype], _IN_KIND[output_tensor.dtype],
            _raw_stream(index),
        )  # fmt: skip
    _lib.raise_for_status(rc)

