"""GPU parity at BASELINE's FULL sizes, where the CPU oracle would take minutes per case: the expected value is computed on the
GPU in fp32 from the library's own dequantised weights -- which tests/test_gpu_parity.py pins bit for bit to the reference's
w_ref and to a plain-torch restatement -- with a plain torch matmul (`x.float() @ W.float()`, TF32 off), i.e. the reference's
`torch.matmul(a, w_ref)` accumulated in fp32.  Tolerance: the north_star's rtol = atol = 1e-2.

  * the reference's own test matrix, largest shape (tests/mixed_precision_gemm_test.py:36: (4096, 2048, 4096), 16 cases), with the
    reference's recipe and its own looser tolerance next to the north_star one;
  * BASELINE config 3 (prefill M = 2048 / 4096, K = 4096, N = 14336) and the README shape 4096 x 8192 x 4096 on the tcgen05 kernel;
  * BASELINE config 4 (W8A16 and per-channel int4, bf16) at Llama sizes, M = 1, 16, 256, 4096;
  * BASELINE config 5's shard shapes (K = 8192, N = 28672 / {1, 2, 4, 8}) at M = 1 and 4096.
"""

from __future__ import annotations

import math

import pytest
import torch

from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import dequantize_weights, mixed_precision_gemm
from helpers import ATOL, RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded(built_lib):
    assert torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False  # the fp32 reference matmul must be fp32
    yield
    torch.backends.cuda.matmul.allow_tf32 = old
    with open("/proc/self/maps") as f:
        assert "libconch_b200.so" in f.read()


def _fp32_reference(x, wq, ws, wzp, st, g):
    w = dequantize_weights(wq, ws, wzp, st.size_bits, st.bias, g)
    return x.float() @ w.float()


def _check(out, ref, what, atol=ATOL):
    torch.testing.assert_close(out.float(), ref, rtol=RTOL, atol=atol, msg=lambda m: f"{what}: {m}")


def _unrounded_w_atol(x, wq, ws, wzp, st, g):
    """The decode kernels never form the rounded W (scale and zero-point are applied per scale group in fp32), the reference
    rounds every weight to the activation dtype first.  The two results differ by sum_k x_k * delta_k, delta_k the rounding
    error of W[k, n]: |delta| <= ulp / 2 with ulp <= 2^-7 |W| for bf16 (2^-10 for fp16), close to uniform and independent, so
    the difference has sigma <= sqrt(K) * rms(x) * 2^-7 * rms(W) / sqrt(12).  With bf16 at K = 4096 that is ~4e-3: five sigma
    exceed the flat atol of 1e-2 on outputs that happen to be near zero, so the bound is the larger of the two."""
    w = dequantize_weights(wq, ws, wzp, st.size_bits, st.bias, g).float()
    ulp_rel = 2.0**-7 if x.dtype == torch.bfloat16 else 2.0**-10
    sigma = math.sqrt(x.shape[1]) * float(x.float().pow(2).mean().sqrt()) * ulp_rel * float(w.pow(2).mean().sqrt()) / math.sqrt(12.0)
    return max(ATOL, 5.0 * sigma)


# ------------------------------------------------------------------------------------------------------------------
# the reference's test matrix, largest shape
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("weight_dtype", ["uint4b8", "uint8b128", "uint4", "uint8"])
@pytest.mark.parametrize("use_zero_points", [True, False])
@pytest.mark.parametrize("input_dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_reference_matrix_largest_shape(weight_dtype, use_zero_points, input_dtype):
    """(4096, 2048, 4096) with the reference test's recipe, quantiser and packer (conch_b200.quant_utils == the oracle ==
    the reference, tests/test_host_logic.py), on the device like the reference test does."""
    m_dim, k_dim, n_dim, g = 4096, 2048, 4096, 128
    st = getattr(quant_utils.scalar_types, weight_dtype)
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = (10 * (torch.rand((m_dim, k_dim), dtype=torch.float32, device="cuda", generator=gen) - 0.3)).to(input_dtype)
    b = (10 * (torch.rand((k_dim, n_dim), dtype=torch.float32, device="cuda", generator=gen) - 0.3)).to(input_dtype)
    w_ref, w_q, w_s, w_zp = quant_utils.quantize_weights(b, st, g, zero_points=use_zero_points)
    w_q_packed = quant_utils.pack_rows(w_q, st.size_bits, *w_q.shape)
    out = mixed_precision_gemm(a, w_q_packed, w_s, w_zp, st.size_bits, st.bias, g)
    # dequantised weights: bit-exact against the quantiser's own w_ref
    assert torch.equal(dequantize_weights(w_q_packed, w_s, w_zp, st.size_bits, st.bias, g), w_ref)
    # the reference test's expression and tolerance (tests/mixed_precision_gemm_test.py:70-77)
    torch.testing.assert_close(out, torch.matmul(a, w_ref), rtol=1e-1, atol=min(5e-2 * math.sqrt(k_dim), 1))
    # north_star tolerance against the fp32-accumulated product.  bf16 outputs of magnitude ~1e3 carry half an ulp (2^-9
    # relative) of output rounding, inside rtol
    _check(out, a.float() @ w_ref.float(), f"{weight_dtype} zp={use_zero_points} {input_dtype}")


# ------------------------------------------------------------------------------------------------------------------
# config 3: prefill on the tcgen05 kernel, full size
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n", [(2048, 4096, 14336), (4096, 4096, 14336), (4096, 8192, 4096)], ids=["ffn_2048", "ffn_4096", "readme"])
def test_prefill_full_size(m, k, n):
    st = quant_utils.scalar_types.uint4
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=3)
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    assert _lib.load().cb200_last_variant() == _lib.TC_TS
    _check(out, _fp32_reference(x, wq, ws, wzp, st, 128), f"{m}x{k}x{n}")


# ------------------------------------------------------------------------------------------------------------------
# config 4: W8A16 and per-channel int4 with bf16 activations, M = 1 .. 4096, Llama sizes
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 16, 256, 4096])
@pytest.mark.parametrize("wname,g", [("uint8", 128), ("uint8b128", 128), ("uint4", -1), ("uint4b8", -1)])
@pytest.mark.parametrize("k,n", [(4096, 4096), (4096, 14336)], ids=["o", "ffn"])
def test_config4_formats_full_size(m, wname, g, k, n):
    st = getattr(quant_utils.scalar_types, wname)
    gs = k if g == -1 else g
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, gs, torch.bfloat16, "cuda", zero_points=not st.has_bias(), seed=7)
    x = torch.randn((m, k), device="cuda", dtype=torch.bfloat16)
    out = mixed_precision_gemm(x, wq, ws, wzp, st.size_bits, st.bias, g)
    fused = _lib.load().cb200_last_variant() in (_lib.DECODE_STREAM, _lib.DECODE_MMA)
    atol = _unrounded_w_atol(x, wq, ws, wzp, st, gs) if fused else ATOL
    _check(out, _fp32_reference(x, wq, ws, wzp, st, gs), f"M={m} {wname} g={g} {k}x{n}", atol)


# ------------------------------------------------------------------------------------------------------------------
# config 5: the 70B FFN shards (K = 8192, N = 28672 / G), decode and prefill
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shards", [1, 2, 4, 8])
@pytest.mark.parametrize("m", [1, 4096])
def test_config5_shard_shapes(shards, m):
    k, n = 8192, 28672 // shards
    st = quant_utils.scalar_types.uint4
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=9)
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    _check(out, _fp32_reference(x, wq, ws, wzp, st, 128), f"M={m} shard 1/{shards}")
