"""GPU parity of the scaled-GEMM sibling path through the C ABI: cb200_scaled_gemm (tcgen05 kind::i8 /
kind::f8f6f4) and cb200_static_scaled_quant, against oracle/scaled_oracle.py and the committed fixtures of the real reference.

  * int8: the accumulation is exact, so the kernel must reproduce the reference OP's arithmetic BIT FOR BIT (scaled_gemm_op);
  * fp8 : the tensor core's fp32 summation order is its own -- rtol = atol = 1e-2 against the fp32 oracle (the reference test
    allows 1e-1, tests/scaled_gemm_test.py:78);
  * quantisers: bit-exact.
"""

from __future__ import annotations

import pathlib

import numpy as np
import pytest
import torch

import scaled_oracle as so
from conch_b200.ops.quantization.fp8 import scaled_fp8_quant
from conch_b200.ops.quantization.gemm import scaled_gemm
from conch_b200.ops.quantization.int8 import scaled_int8_quant
from test_oracle_scaled import GEMM_FILES, OD, QUANT_FILES, _load_gemm, _t16

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded(built_lib):
    assert torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    yield
    with open("/proc/self/maps") as f:
        assert "libconch_b200.so" in f.read()


def _cuda(*ts):
    return [None if t is None else t.cuda() for t in ts]


@pytest.mark.parametrize("path", GEMM_FILES, ids=lambda p: p.stem)
def test_golden_fixtures_of_the_real_reference(path):
    a, b, sa, sb, bias, od, ref = _load_gemm(path)
    ad, bd, sad, sbd, biasd = _cuda(a, b, sa, sb, bias)
    out = scaled_gemm(ad, bd, sad, sbd, od, biasd)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-1, atol=1e-1)  # the reference test's tolerance
    if a.dtype == torch.int8:
        assert torch.equal(out.cpu().view(torch.int16), so.scaled_gemm_op(a, b, sa, sb, od, bias).view(torch.int16))
    else:
        torch.testing.assert_close(out.cpu().float(), so.scaled_gemm_op(a, b, sa, sb, od, bias).float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize(("m", "k", "n"), [(128, 256, 128), (1024, 1024, 1024), (4096, 2048, 4096), (1, 4096, 4096), (77, 208, 136), (300, 1040, 72)])
@pytest.mark.parametrize("kind", ["int8", "fp8"])
@pytest.mark.parametrize("od", [torch.float16, torch.bfloat16, torch.float32], ids=["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("scalars", [(True, True), (False, False), (True, False)], ids=["ss", "vv", "sv"])
@pytest.mark.parametrize("use_bias", [True, False], ids=["bias", "nobias"])
def test_reference_matrix_and_ragged_shapes(m, k, n, kind, od, scalars, use_bias):
    """The reference test's three shapes (tests/scaled_gemm_test.py:31) plus decode-like and ragged ones (M, N not multiples of
    the 128-wide tile; K not a multiple of the 128-byte block)."""
    a, b, sa, sb, bias = so.make_scaled_problem(m, k, n, kind, od, scalars[0], scalars[1], use_bias, seed=m + n)
    ad, bd, sad, sbd, biasd = _cuda(a, b, sa, sb, bias)
    out = scaled_gemm(ad, bd, sad, sbd, od, biasd, strict=True)
    assert out.dtype == od and tuple(out.shape) == (m, n)
    # expected value on the GPU in fp32 (exact for int8: |acc| < 2^24), the op's epilogue arithmetic on top
    acc = ad.float() @ bd.float()
    want = (acc * (sad.reshape(-1, 1).float() * sbd.reshape(1, -1).float())).to(od)
    if use_bias:
        want = want + biasd
    if kind == "int8":
        assert torch.equal(out, want)
    else:
        torch.testing.assert_close(out.float(), want.float(), rtol=1e-2, atol=1e-2)


def test_row_major_b_is_transposed_once_and_arguments_are_checked():
    a, b, sa, sb, _ = so.make_scaled_problem(64, 256, 128, "int8", torch.float16, False, False, False, seed=1)
    ad, bd, sad, sbd = _cuda(a, b, sa, sb)
    want = scaled_gemm(ad, bd, sad, sbd, torch.float16)
    assert torch.equal(scaled_gemm(ad, bd.contiguous(), sad, sbd, torch.float16), want)  # [K, N] row-major
    with pytest.raises(ValueError):
        scaled_gemm(ad, bd.to(torch.float16).to(so.FP8), sad, sbd, torch.float16, strict=True)  # a / b dtypes differ
    with pytest.raises(ValueError):
        scaled_gemm(ad, bd, sad[:5], sbd, torch.float16)
    with pytest.raises(NotImplementedError):
        scaled_gemm(ad.to(torch.float16), bd.to(torch.float16), sad, sbd, torch.float16)
    with pytest.raises(ValueError):  # K % 16 != 0
        scaled_gemm(ad[:, :250].contiguous(), bd[:250].contiguous(), sad, sbd, torch.float16)


@pytest.mark.parametrize("path", QUANT_FILES, ids=lambda p: p.stem)
def test_quantiser_golden_fixtures(path):
    z = np.load(path)
    dt = OD[path.stem.split("_")[2]]
    x8 = torch.from_numpy(z["x8"]) if dt == torch.float32 else _t16(z["x8"], dt)
    xf = torch.from_numpy(z["xf"]) if dt == torch.float32 else _t16(z["xf"], dt)
    s = torch.from_numpy(z["scale"]).cuda()
    q8, _ = scaled_int8_quant(x8.cuda(), s)
    qf, _ = scaled_fp8_quant(xf.cuda(), s)
    assert np.array_equal(q8.cpu().view(torch.uint8).numpy(), z["q8"])
    assert np.array_equal(qf.cpu().view(torch.uint8).numpy(), z["qf"])


@pytest.mark.parametrize("num_tokens", [1, 7, 83, 4096])
@pytest.mark.parametrize("hidden_size", [16, 67, 768, 5137, 8193])
@pytest.mark.parametrize("dtype", [torch.half, torch.bfloat16, torch.float], ids=["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("scale", [0.1, 2.1])
def test_quantisers_reference_matrix(num_tokens, hidden_size, dtype, scale):
    """The reference tests' matrix (tests/int8_quant_kernels_test.py:15-19, fp8_quant_kernels_test.py) -- bit-exact here, where
    the reference allows one unit (int8) / dequantised closeness (fp8)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    s = torch.tensor([scale], dtype=torch.float32, device="cuda")
    x = torch.rand((num_tokens, hidden_size), dtype=dtype, device="cuda", generator=g) * 1000 - 200
    q8, s_out = scaled_int8_quant(x, s)
    assert s_out is s and q8.dtype == torch.int8 and q8.shape == x.shape
    assert torch.equal(q8, so.scaled_int8_quant(x, s))
    xf = torch.rand((num_tokens, hidden_size), dtype=dtype, device="cuda", generator=g) * 6 - 2
    qf, _ = scaled_fp8_quant(xf, s)
    assert torch.equal(qf.view(torch.uint8), so.scaled_fp8_quant(xf, s).view(torch.uint8))
    with pytest.raises(NotImplementedError):
        scaled_int8_quant(x)
    with pytest.raises(NotImplementedError):
        scaled_fp8_quant(x)
