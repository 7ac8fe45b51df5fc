"""Public op: `mixed_precision_gemm` with the reference's exact Python signature.

Drop-in for `conch.ops.quantization.gemm.mixed_precision_gemm` (reference conch/ops/quantization/gemm.py:133-168):
same positional order, keyword names and defaults, same packing layout (conch/third_party/vllm/quant_utils.py:104-125),
same group_size semantics, same exception types in strict mode.  Differences, all supersets:
  * `group_size=-1` is accepted as an alias of K (the reference op would read out of bounds);
  * shapes / dtypes are always validated (cheaply) -- the reference only checks when `strict=True`;
  * `w_zp` may be int32 (reference quantiser), uint8 (what vLLM passes) or the activation dtype;
  * the output is allocated with `torch.empty` (the reference memsets with `torch.zeros`, gemm.py:164).
There is no CPU / Triton fallback: inputs must live on a B200.
"""

from __future__ import annotations

import ctypes
import weakref
from collections.abc import Sequence
from typing import Final

import torch

from conch_b200 import _lib
from conch_b200.kernels.quantization import gemm as _kernels_gemm
from conch_b200.kernels.quantization.gemm import (
    ChannelScaleMode,
    dequantize_launcher,
    mixed_precision_gemm_launcher,
    MixedPrecisionMatmulMetadata,
    scaled_gemm_launcher,
    ScaledMatmulMetadata,
    WeightGroupMode,
)
from conch_b200.platform import current_platform

_ACT_DTYPES: Final = (torch.float16, torch.bfloat16)
_OUT_DTYPES: Final = (torch.float16, torch.bfloat16, torch.float32)


def _check_contiguous(x: torch.Tensor, w_q_packed: torch.Tensor, w_s: torch.Tensor, w_zp: torch.Tensor | None) -> bool:
    return (
        x.is_contiguous()
        and w_q_packed.is_contiguous()
        and w_s.is_contiguous()
        and (w_zp is None or w_zp.is_contiguous())
    )


def _deduce_weight_group_mode(w_zp: torch.Tensor | None) -> WeightGroupMode:
    # reference conch/ops/quantization/gemm.py:34-38
    return WeightGroupMode.SYMMETRIC_NO_SHIFT if w_zp is None else WeightGroupMode.SYMMETRIC_WITH_SHIFT


def create_mixed_precision_metadata(
    x: torch.Tensor,
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    weight_size_bits: int,
    weight_bias: int,
    group_size: int,
    *,
    output_dtype: torch.dtype | None = None,
    acc_dtype: torch.dtype | None = None,
    meta_dtype: torch.dtype | None = None,
    scaled_activations: bool = False,
    strict: bool = False,
) -> MixedPrecisionMatmulMetadata:
    """Verify sizes / dtypes and deduce metadata (reference conch/ops/quantization/gemm.py:41-130).

    Runs on every call of the op (the reference only validates when `strict`), so it is written for speed: plain
    attribute reads, no temporaries -- about 3 us.
    """
    zero_is_scalar = w_zp is not None and w_zp.numel() == 1  # :92
    if x.dim() != 2:
        raise ValueError(f"Unexpected number of dimensions of input tensor x: {x.dim()}")
    if w_q_packed.dim() != 2:
        raise ValueError(f"Unexpected number of dimensions of input tensor w_q_packed: {w_q_packed.dim()}")
    if w_s.dim() != 2:
        raise ValueError(f"Unexpected number of dimensions of input tensor w_s: {w_s.dim()}")
    # a scalar zero-point may have any rank unless strict (reference :71-73 checks rank only)
    if w_zp is not None and w_zp.dim() != 2 and (strict or not zero_is_scalar):
        raise ValueError(f"Unexpected number of dimensions of input tensor w_zp: {w_zp.dim()}")
    wq_dtype = w_q_packed.dtype
    if wq_dtype is not torch.int32 and wq_dtype is not torch.uint32:
        raise ValueError(f"Invalid datatype for packed weights: {wq_dtype}")
    if weight_size_bits != 4 and weight_size_bits != 8:
        raise NotImplementedError(f"weight_size_bits={weight_size_bits}: only 4 and 8 bit weights are implemented")
    x_dtype = x.dtype
    if x_dtype is not torch.float16 and x_dtype is not torch.bfloat16:
        raise NotImplementedError(f"activation dtype {x_dtype}: only float16 / bfloat16 are implemented")

    elements_per_sample = 32 // weight_size_bits  # :83-84
    m_dim, k_dim = x.shape  # :86-87
    packed_rows, n_dim = w_q_packed.shape
    if group_size == -1:
        group_size = k_dim
    if packed_rows * elements_per_sample != k_dim:
        raise ValueError(
            f"w_q_packed has {packed_rows} rows of {elements_per_sample} values, x has K={k_dim} columns"
        )
    unpack_mask = 2**weight_size_bits - 1  # :89

    if group_size <= 0 or k_dim % group_size:
        raise ValueError(f"group_size={group_size} must divide K={k_dim}")
    meta_rows = k_dim // group_size  # :96,101
    s_shape = w_s.shape
    if s_shape[0] != meta_rows or s_shape[1] != n_dim:
        raise ValueError(f"Invalid w_s shape (expected: {(meta_rows, n_dim)}, actual: {tuple(s_shape)})")
    if w_zp is not None and not zero_is_scalar:
        z_shape = w_zp.shape
        if z_shape[0] != meta_rows or z_shape[1] != n_dim:
            raise ValueError(f"Invalid w_zp shape (expected: {(meta_rows, n_dim)}, actual: {tuple(z_shape)})")
    if strict and scaled_activations:  # :108-111 (ignored when not strict, like the reference)
        raise NotImplementedError("Scaled activations not yet implemented (need to deduce correct channel_scale_mode)")

    if w_s.dtype is not x_dtype:
        raise NotImplementedError(f"w_s dtype {w_s.dtype} must equal the activation dtype {x_dtype}")
    out_dt = x_dtype if output_dtype is None else output_dtype
    if out_dt not in _OUT_DTYPES:
        raise NotImplementedError(f"output_dtype {out_dt} not implemented")
    acc_dt = torch.float32 if acc_dtype is None else acc_dtype
    if acc_dt is not torch.float32:
        raise NotImplementedError("only float32 accumulation is implemented (the reference default)")
    meta_dt = x_dtype if meta_dtype is None else meta_dtype
    if meta_dt is not x_dtype:
        raise NotImplementedError("meta_dtype must equal the activation dtype (the reference default)")

    return MixedPrecisionMatmulMetadata(
        m_dim=m_dim,
        k_dim=k_dim,
        n_dim=n_dim,
        weight_size_bits=weight_size_bits,
        weight_bias=weight_bias,
        group_size=group_size,
        elements_per_sample=elements_per_sample,
        zero_is_scalar=zero_is_scalar,
        unpack_mask=unpack_mask,
        data_contiguous=_check_contiguous(x, w_q_packed, w_s, w_zp),
        input_dtype=x_dtype,
        output_dtype=out_dt,
        acc_dtype=acc_dt,
        meta_dtype=meta_dt,
        channel_scale_mode=ChannelScaleMode.NONE,
        weight_group_mode=WeightGroupMode.SYMMETRIC_NO_SHIFT if w_zp is None else WeightGroupMode.SYMMETRIC_WITH_SHIFT,  # :34-38
    )


_checked_devices: set[int | None] = set()  # device indices already verified to be sm_100 (the check costs ~3 us)


def _require_device(*tensors: torch.Tensor | None) -> None:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "conch_b200.mixed_precision_gemm needs CUDA tensors on a B200: there is no CPU fallback "
                f"(got a tensor on {t.device})"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"all tensors must be on the same device ({dev} vs {t.device})")
    index = dev.index if dev is not None else None
    if index not in _checked_devices:
        current_platform.check(index)
        _checked_devices.add(index)


def mixed_precision_gemm(
    x: torch.Tensor,
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    weight_size_bits: int,
    weight_bias: int,
    group_size: int,
    *,
    output_dtype: torch.dtype | None = None,
    acc_dtype: torch.dtype | None = None,
    meta_dtype: torch.dtype | None = None,
    scaled_activations: bool = False,
    strict: bool = False,
    variant: int = _lib.AUTO,
    out: torch.Tensor | None = None,
    peer_out_ptrs: Sequence[int] | None = None,
    signal: tuple[Sequence[int], int] | None = None,
    x_ready: tuple[int, int, int] | None = None,
    bias: torch.Tensor | None = None,
) -> torch.Tensor:
    """Mixed precision GEMM: `x[M,K] @ dequant(w_q_packed[K/pf,N], w_s, w_zp) -> [M,N]`.

    Signature of the reference op (conch/ops/quantization/gemm.py:133-147) plus optional keywords:
    `variant` forces a kernel family (debug / benchmarking), `out` reuses a caller-owned output buffer, `peer_out_ptrs`
    (multi-GPU, decode regime: M <= 8) lists peer-GPU buffers the kernel also stores the result to, `signal` = (peer flag
    pointers, local flag pointer) makes it bump completion counters on every GPU afterwards (conch_b200/sharded.py), `x_ready` =
    (flags pointer, count, expected pointer) declares `x` a gather buffer other GPUs push into, complete once those counters have
    been raised -- the kernel waits for them itself while its weights stream in (`cb200_mpgemm_chain`) -- and
    `bias` ([N], output dtype) is added in the kernels' epilogue -- to the fp32 sum, before the one rounding -- instead of by
    the separate `output.add_(bias)` kernel a caller of the reference op runs (the reference's sibling op does exactly that,
    conch/ops/quantization/gemm.py:249-250).
    """
    # ---- repeat calls on the same weight tensors (a decode loop): everything that depends on the weights alone was validated
    # and resolved by the first call; only x / out are looked at again.  4-5 us of host time instead of 11 -- at M=1 the GEMVs of
    # a Llama-3-8B token take 9.7 us each on the device, so the full path is what an eager caller would wait for.
    # (peer pointers ride along when the caller hands them over as a ready C array: conch_b200/sharded.py pushed())
    plain = (output_dtype is None and bias is None and signal is None and x_ready is None and acc_dtype is None
             and meta_dtype is None and not strict and (peer_out_ptrs is None or isinstance(peer_out_ptrs, ctypes.Array)))  # fmt: skip
    if plain:
        e = _fast_entries.get(id(w_q_packed))
        if (e is not None and e[0]() is w_q_packed and e[1] is w_s and e[2] is w_zp and e[3] == weight_size_bits
                and e[4] == weight_bias and e[5] == group_size and x.dtype is e[6] and x.dim() == 2 and x.is_contiguous()
                and x.device == e[7]):  # fmt: skip
            m_dim, k_dim = x.shape
            if k_dim == e[8] and m_dim > 0:
                if out is None:
                    out = torch.empty((m_dim, e[9]), device=e[7], dtype=e[6])
                elif out.dim() != 2 or out.shape[0] != m_dim or out.shape[1] != e[9] or out.dtype is not e[6] or not out.is_contiguous():
                    raise ValueError("`out` must be a contiguous [M, N] tensor of the output dtype")
                if _fast_launch(e, x, out, m_dim, variant, peer_out_ptrs):
                    return out
    _require_device(x, w_q_packed, w_s, w_zp)
    metadata = create_mixed_precision_metadata(
        x, w_q_packed, w_s, w_zp, weight_size_bits, weight_bias, group_size,
        output_dtype=output_dtype, acc_dtype=acc_dtype, meta_dtype=meta_dtype,
        scaled_activations=scaled_activations, strict=strict,
    )  # fmt: skip
    register = plain and metadata.data_contiguous and metadata.m_dim > 0
    if not metadata.data_contiguous:
        # the reference honours arbitrary strides (gemm.py:518-525); the CUDA kernels want dense row-major
        x, w_q_packed, w_s = x.contiguous(), w_q_packed.contiguous(), w_s.contiguous()
        w_zp = None if w_zp is None else w_zp.contiguous()
    if out is None:
        out = torch.empty((metadata.m_dim, metadata.n_dim), device=x.device, dtype=metadata.output_dtype)
    elif tuple(out.shape) != (metadata.m_dim, metadata.n_dim) or out.dtype != metadata.output_dtype or not out.is_contiguous():
        raise ValueError("`out` must be a contiguous [M, N] tensor of the output dtype")
    if bias is not None:
        if tuple(bias.shape) != (metadata.n_dim,) or bias.dtype != metadata.output_dtype or bias.device != x.device:
            raise ValueError(f"`bias` must be a [{metadata.n_dim}] tensor of the output dtype {metadata.output_dtype} on {x.device}")
        bias = bias.contiguous()
    if metadata.m_dim == 0:
        return out
    try:
        mixed_precision_gemm_launcher(
            out, x, w_q_packed, w_s, w_zp, metadata, variant=variant, peer_out_ptrs=peer_out_ptrs, signal=signal, x_ready=x_ready, bias=bias
        )
    except ValueError as e:
        # the kernels read with 16-byte vectors / TMA: a dense view that starts at an odd element of a bigger buffer is copied
        # and the call repeated (the reference's Triton kernel accepts it as is).  Checked here, on the failure path, so that the
        # aligned case pays nothing for it.
        if "16-byte aligned" not in str(e):
            raise
        mixed_precision_gemm_launcher(
            out, _dense_aligned(x), _dense_aligned(w_q_packed), _dense_aligned(w_s), None if w_zp is None else _dense_aligned(w_zp),
            metadata, variant=variant, peer_out_ptrs=peer_out_ptrs, signal=signal, x_ready=x_ready, bias=bias,
        )  # fmt: skip
        register = False
    if register:
        _fast_register(x, w_q_packed, w_s, w_zp, metadata, group_size)
    return out


# id(w_q_packed) -> (weakref(w_q_packed), w_s, w_zp, bits, bias, group_size as passed, dtype, device, K, N, resolved group size,
#                    zero-point kind, zero_is_scalar, dtype code, device index); dropped when the packed tensor dies
_fast_entries: dict[int, tuple] = {}


def _fast_register(x, w_q_packed, w_s, w_zp, metadata: MixedPrecisionMatmulMetadata, group_size_arg: int) -> None:
    from conch_b200.kernels.quantization.gemm import _DTYPE_CODE, zero_point_kind

    key = id(w_q_packed)
    try:
        ref = weakref.ref(w_q_packed, lambda _r, key=key: _fast_entries.pop(key, None))
    except TypeError:  # (a tensor subclass without weak references: no fast path)
        return
    if len(_fast_entries) > 4096:
        _fast_entries.clear()
    device = x.device
    _fast_entries[key] = (
        ref, w_s, w_zp, metadata.weight_size_bits, metadata.weight_bias, group_size_arg, metadata.input_dtype, device,
        metadata.k_dim, metadata.n_dim, metadata.group_size, zero_point_kind(x, w_zp), int(metadata.zero_is_scalar),
        _DTYPE_CODE[metadata.input_dtype], device.index if device.index is not None else torch.cuda.current_device(),
    )  # fmt: skip


def _fast_launch(e: tuple, x: torch.Tensor, out: torch.Tensor, m_dim: int, variant: int, peers=None) -> bool:
    """The launcher without its bookkeeping; False sends the call down the full path (which also words the errors)."""
    kg = _kernels_gemm
    index = e[14]
    if index != torch.cuda.current_device():
        return False
    lib = _lib.load()
    stream = kg._raw_stream(index)
    ws = kg._workspaces.get((index, stream))
    if ws is None:
        return False
    w_zp = e[2]
    code = e[13]
    if peers is None:
        rc = lib.cb200_mpgemm(
            x.data_ptr(), e[0]().data_ptr(), e[1].data_ptr(), None if w_zp is None else w_zp.data_ptr(), out.data_ptr(),
            m_dim, e[8], e[9], e[3], e[4], e[10], code, code, e[11], e[12], ws.data_ptr(), ws.numel(), variant, stream,
        )  # fmt: skip
    else:
        rc = lib.cb200_mpgemm_push(
            x.data_ptr(), e[0]().data_ptr(), e[1].data_ptr(), None if w_zp is None else w_zp.data_ptr(), out.data_ptr(),
            m_dim, e[8], e[9], e[3], e[4], e[10], code, code, e[11], e[12], ws.data_ptr(), ws.numel(), variant, stream,
            peers, len(peers),
        )  # fmt: skip
    return rc == 0


class PreparedWeights:
    """One layer's quantised weights, prepared ONCE for repeated calls (the vLLM-style caller).

    The stateless op re-validates and re-derives everything per call and streams whatever zero-point dtype the quantiser
    produced.  Preparation
      * narrows per-group integer zero-points to uint8 (the reference quantiser emits int32, i.e. 4 bytes per group and
        column where 4 or 8 bits carry the information: -3 bytes per 64 packed int4 bytes, 4.6 % less HBM traffic per GEMV),
      * makes every tensor dense, 16-byte aligned and resident on one device,
      * resolves the metadata (shapes, group size, dtypes) once,
    and `__call__(x, bias=None, out=None)` runs the same kernels as `mixed_precision_gemm` with the bias fused into their
    epilogue.  Results are bit-identical to the stateless op on the same inputs (the kernels read uint8 or int32 zero-points
    through the same arithmetic)."""

    def __init__(self, w_q_packed: torch.Tensor, w_s: torch.Tensor, w_zp: torch.Tensor | None, weight_size_bits: int,
                 weight_bias: int, group_size: int) -> None:  # fmt: skip
        _require_device(w_q_packed, w_s, w_zp)
        if weight_size_bits not in (4, 8):
            raise NotImplementedError(f"weight_size_bits={weight_size_bits}")
        self.weight_size_bits, self.weight_bias = weight_size_bits, weight_bias
        pf = 32 // weight_size_bits
        self.k_dim, self.n_dim = w_q_packed.shape[0] * pf, w_q_packed.shape[1]
        self.group_size = self.k_dim if group_size == -1 else group_size
        if tuple(w_s.shape) != (self.k_dim // self.group_size, self.n_dim):
            raise ValueError(f"Invalid w_s shape {tuple(w_s.shape)}")
        self.w_q_packed = _dense_aligned(w_q_packed)
        self.w_s = _dense_aligned(w_s)
        if w_zp is not None and w_zp.numel() > 1 and w_zp.dtype in (torch.int32, torch.int64, torch.int16):
            lo, hi = int(w_zp.min()), int(w_zp.max())
            if lo < 0 or hi > 255:
                raise ValueError(f"zero-points outside [0, 255] ({lo}..{hi}) cannot be narrowed to uint8")
            w_zp = w_zp.to(torch.uint8)
        self.w_zp = None if w_zp is None else _dense_aligned(w_zp)

    def __call__(self, x: torch.Tensor, bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
                 output_dtype: torch.dtype | None = None, variant: int = _lib.AUTO) -> torch.Tensor:  # fmt: skip
        return mixed_precision_gemm(
            x, self.w_q_packed, self.w_s, self.w_zp, self.weight_size_bits, self.weight_bias, self.group_size,
            output_dtype=output_dtype, out=out, bias=bias, variant=variant,
        )  # fmt: skip


def _dense_aligned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def prepare_weights(w_q_packed: torch.Tensor, w_s: torch.Tensor, w_zp: torch.Tensor | None, weight_size_bits: int,
                    weight_bias: int, group_size: int) -> PreparedWeights:  # fmt: skip
    """Prepare one layer's weights for repeated `mixed_precision_gemm` calls; see PreparedWeights."""
    return PreparedWeights(w_q_packed, w_s, w_zp, weight_size_bits, weight_bias, group_size)


def dequantize_weights(
    w_q_packed: torch.Tensor,
    w_s: torch.Tensor,
    w_zp: torch.Tensor | None,
    weight_size_bits: int,
    weight_bias: int,
    group_size: int,
) -> torch.Tensor:
    """W[K,N] in `w_s.dtype`, bit-identical to the reference quantiser's `w_ref`
    (conch/third_party/vllm/quant_utils.py:74) -- the `_dequantize` stage run on its own."""
    _require_device(w_q_packed, w_s, w_zp)
    pf = 32 // weight_size_bits
    k_dim, n_dim = w_q_packed.shape[0] * pf, w_q_packed.shape[1]
    if group_size == -1:
        group_size = k_dim
    if w_s.dtype not in _ACT_DTYPES:
        raise NotImplementedError(f"scale dtype {w_s.dtype}")
    if tuple(w_s.shape) != (k_dim // group_size, n_dim):
        raise ValueError(f"Invalid w_s shape {tuple(w_s.shape)}")
    out = torch.empty((k_dim, n_dim), device=w_q_packed.device, dtype=w_s.dtype)
    dequantize_launcher(
        out, w_q_packed.contiguous(), w_s.contiguous(), None if w_zp is None else w_zp.contiguous(),
        weight_size_bits, weight_bias, group_size,
    )  # fmt: skip
    return out


def create_scaled_metadata(
    a: torch.Tensor,
    b: torch.Tensor,
    scale_a: torch.Tensor,
    scale_b: torch.Tensor,
    output_dtype: torch.dtype,
    strict: bool = False,
) -> ScaledMatmulMetadata:
    """Verify sizes and dtypes of tensors and deduce metadata parameters (reference conch/ops/quantization/gemm.py:171-230:
    same checks, same messages, same result)."""
    if strict:
        expected_input_matrix_rank = 2
        if (a_rank := len(a.shape)) != expected_input_matrix_rank:
            raise ValueError(f"Unexpected number of dimensions of input tensor a: {a_rank}")
        if (b_rank := len(b.shape)) != expected_input_matrix_rank:
            raise ValueError(f"Unexpected number of dimensions of input tensor b: {b_rank}")
        if a.dtype != b.dtype:
            raise ValueError(f"Input tensors a and b must have the same datatype (a: {a.dtype}, b: {b.dtype})")
    m_dim, k_dim = a.shape
    _, n_dim = b.shape
    if strict:
        if scale_a.numel() != 1:
            if (scale_a_rank := len(scale_a.shape)) != expected_input_matrix_rank:
                raise ValueError(f"Unexpected number of dimensions of input tensor scale_a: {scale_a_rank}")
            if scale_a.shape[0] != m_dim:
                raise ValueError(f"Invalid scale_a shape (expected: ({m_dim},), actual: {scale_a.shape})")
        if scale_b.numel() != 1:
            if (scale_b_rank := len(scale_b.shape)) != expected_input_matrix_rank:
                raise ValueError(f"Unexpected number of dimensions of input tensor scale_b: {scale_b_rank}")
            if scale_b.shape[0] != n_dim:
                raise ValueError(f"Invalid scale_b shape (expected: ({n_dim},), actual: {scale_b.shape})")
    return ScaledMatmulMetadata(
        m_dim=m_dim,
        k_dim=k_dim,
        n_dim=n_dim,
        data_contiguous=(a.is_contiguous() and b.is_contiguous() and scale_a.is_contiguous() and scale_b.is_contiguous()),
        input_dtype=a.dtype,
        output_dtype=output_dtype,
        acc_dtype=torch.float32 if a.is_floating_point() else torch.int32,
        meta_dtype=scale_a.dtype,
        channel_scale_mode=ChannelScaleMode.WEIGHT_AND_ACTIVATION,
        weight_group_mode=WeightGroupMode.NONE,
    )


def scaled_gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    scale_a: torch.Tensor,
    scale_b: torch.Tensor,
    output_dtype: torch.dtype,
    bias: torch.Tensor | None = None,
    strict: bool = False,
) -> torch.Tensor:
    """Scaled GEMM: `cast((a[M,K] @ b[K,N]) * (scale_a[M,1] * scale_b[N,1]^T)) (+ bias[N])` for int8 / float8_e4m3fn operands.

    Signature of the reference op (conch/ops/quantization/gemm.py:233-252).  The bias goes into the kernel's epilogue, which
    reproduces the reference's two-kernel arithmetic (round, add, round)."""
    _require_device(a, b, scale_a, scale_b, bias)
    metadata = create_scaled_metadata(a, b, scale_a, scale_b, output_dtype, strict=strict)
    if a.dtype != b.dtype:
        raise ValueError(f"Input tensors a and b must have the same datatype (a: {a.dtype}, b: {b.dtype})")
    if scale_a.numel() not in (1, metadata.m_dim) or scale_b.numel() not in (1, metadata.n_dim):
        raise ValueError(f"scale_a / scale_b must have 1 or M / N elements (got {scale_a.numel()}, {scale_b.numel()})")
    if bias is not None:
        if bias.numel() != metadata.n_dim:
            raise ValueError(f"`bias` must have N = {metadata.n_dim} elements")
        bias = bias.to(output_dtype).contiguous().reshape(-1)
    output = torch.empty((metadata.m_dim, metadata.n_dim), device=a.device, dtype=output_dtype)
    if metadata.m_dim == 0 or metadata.n_dim == 0:
        return output
    scaled_gemm_launcher(output, a, b, scale_a, scale_b, metadata, bias=bias)
    return output

