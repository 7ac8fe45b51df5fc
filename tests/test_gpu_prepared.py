"""SURVEY.md 8f row 1 -- prepared weights and the bias fused into the kernels' epilogue -- through the C ABI (cb200_mpgemm_bias).

The reference's caller runs `output.add_(bias)` as a second kernel (its sibling op does exactly that,
conch/ops/quantization/gemm.py:249-250): out = RN(RN(acc) + bias).  The fused epilogue computes RN(acc + bias) -- one rounding
less -- so it is checked against the fp32 oracle + bias at the north_star tolerance, and against the two-kernel form within
one unit in the last place of the output dtype.
"""

from __future__ import annotations

import pytest
import torch

import mpgemm_oracle as oracle
from conch_b200 import _lib
from conch_b200.ops.quantization.gemm import mixed_precision_gemm, prepare_weights
from helpers import assert_gemm_close, bits, to_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded(built_lib):
    assert torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    yield
    with open("/proc/self/maps") as f:
        assert "libconch_b200.so" in f.read()


CASES = [
    # (M, K, N, weights, dtype, variant): every kernel family's store path, split-K tails included
    (1, 1024, 512, "uint4", torch.float16, _lib.AUTO),
    (5, 2048, 320, "uint4b8", torch.bfloat16, _lib.DECODE_STREAM),
    (16, 1024, 256, "uint8", torch.float16, _lib.DECODE_STREAM),
    (3, 1024, 512, "uint8", torch.bfloat16, _lib.DECODE_MMA),
    (12, 2048, 256, "uint4", torch.float16, _lib.DECODE_MMA),
    (64, 1024, 512, "uint4", torch.float16, _lib.TC_TS),
    (200, 4096, 256, "uint4", torch.float16, _lib.TC_TS),  # few tiles, long K: the split-K tail + its reduce kernel
    (160, 512, 256, "uint8b128", torch.bfloat16, _lib.TC_SS),
    (300, 1024, 384, "uint4", torch.float16, _lib.TC2_TS),
]


@pytest.mark.parametrize("m,k,n,wname,dtype,variant", CASES)
@pytest.mark.parametrize("out_dtype", [None, torch.float32], ids=["act", "f32"])
def test_fused_bias(m, k, n, wname, dtype, variant, out_dtype):
    zp = wname in ("uint4", "uint8")
    p = oracle.make_problem(m, k, n, wname, 128, dtype, zp, seed=31)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    od = out_dtype or dtype
    bias = (torch.randn((n,), generator=torch.Generator().manual_seed(5)) * 0.5).to(od)
    got = mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, 128, variant=variant, output_dtype=od, bias=bias.cuda())
    assert got.dtype == od
    want32 = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, 128, output_dtype=torch.float32)
    assert_gemm_close(got, want32.float() + bias.float(), f"{m}x{k}x{n} {wname} bias")
    # against the reference caller's two kernels: RN(RN(acc) + bias) vs RN(acc + bias) -- at most one ulp of the output apart
    plain = mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, 128, variant=variant, output_dtype=od)
    two = plain.clone().add_(bias.cuda())
    ulp = {torch.float16: 2.0**-10, torch.bfloat16: 2.0**-7, torch.float32: 2.0**-23}[od]
    scale = torch.maximum(two.float().abs(), plain.float().abs()).clamp_min(2.0**-14)
    assert bool(((got.float() - two.float()).abs() <= 2.0 * ulp * scale).all())


def test_bias_argument_checks():
    p = oracle.make_problem(2, 512, 256, "uint4", 128, torch.float16, True, seed=1)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    with pytest.raises(ValueError):
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, bias=torch.zeros(255, device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, bias=torch.zeros(256, device="cuda", dtype=torch.float32))
    with pytest.raises(NotImplementedError):
        peer = torch.zeros((2, 256), device="cuda", dtype=torch.float16)
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, bias=torch.zeros(256, device="cuda", dtype=torch.float16), peer_out_ptrs=[peer.data_ptr()])


@pytest.mark.parametrize("m", [1, 8, 16, 256])
@pytest.mark.parametrize("wname,dtype", [("uint4", torch.float16), ("uint8", torch.float16), ("uint4", torch.bfloat16)])
def test_prepared_weights_equal_the_stateless_op_bit_for_bit(m, wname, dtype):
    """prepare_weights narrows the int32 zero-points to uint8 and fixes layout / alignment; the kernels read either through the
    same arithmetic, so nothing may change in the result."""
    p = oracle.make_problem(m, 2048, 1024, wname, 128, dtype, True, seed=13)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    assert wzp.dtype == torch.int32
    # a dense view at an odd offset of a bigger buffer (the stateless op copies it)
    flat = torch.empty(ws.numel() + 3, device="cuda", dtype=ws.dtype)
    ws_odd = flat[3:].view_as(ws).copy_(ws)
    assert ws_odd.data_ptr() % 16 != 0
    stateless = mixed_precision_gemm(x, wq, ws_odd, wzp, p.wtype.bits, p.wtype.bias, 128)
    prepared = prepare_weights(wq, ws_odd, wzp, p.wtype.bits, p.wtype.bias, 128)
    assert prepared.w_zp.dtype == torch.uint8 and prepared.w_s.data_ptr() % 16 == 0
    got = prepared(x)
    assert torch.equal(bits(got), bits(stateless))
    assert_gemm_close(got, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, 128))
    bias = torch.randn((1024,), device="cuda").to(dtype)
    assert torch.equal(bits(prepared(x, bias=bias)), bits(mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, 128, bias=bias)))
