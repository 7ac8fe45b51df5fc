"""The `conch` shim (shim/conch): the reference's import paths for the hot path resolve to conch_b200, with the call signatures
the reference's own test uses (reference tests/mixed_precision_gemm_test.py:12-16, 27-33, 72-74).  CPU only: no kernel runs."""

from __future__ import annotations

import importlib
import inspect
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _with_shim():
    sys.path.insert(0, str(ROOT / "shim"))
    for name in [m for m in sys.modules if m == "conch" or m.startswith("conch.")]:
        del sys.modules[name]


def test_shim_modules_resolve_to_conch_b200():
    _with_shim()
    try:
        gemm = importlib.import_module("conch.ops.quantization.gemm")
        plat = importlib.import_module("conch.platforms")
        qu = importlib.import_module("conch.third_party.vllm.quant_utils")
        stt = importlib.import_module("conch.third_party.vllm.scalar_type")
        ut = importlib.import_module("conch.third_party.vllm.utils")
        import conch_b200.ops.quantization.gemm as ours
        import conch_b200.quant_utils as ours_q

        assert gemm.mixed_precision_gemm is ours.mixed_precision_gemm
        assert gemm.create_mixed_precision_metadata is ours.create_mixed_precision_metadata
        assert qu.quantize_weights is ours_q.quantize_weights and qu.pack_rows is ours_q.pack_rows
        assert stt.scalar_types.uint4b8.bias == 8 and stt.scalar_types.uint8b128.size_bits == 8
        assert isinstance(stt.scalar_types.uint4, stt.ScalarType)
        assert plat.current_platform.device == "cuda"
        ut.seed_everything(0)
        a = torch.rand(3)
        ut.seed_everything(0)
        assert torch.equal(a, torch.rand(3))
    finally:
        sys.path.remove(str(ROOT / "shim"))
        for name in [m for m in sys.modules if m == "conch" or m.startswith("conch.")]:
            del sys.modules[name]


def test_signatures_the_reference_test_uses():
    from conch_b200.ops.quantization.gemm import mixed_precision_gemm
    from conch_b200.quant_utils import pack_rows, quantize_weights, scalar_types

    # mixed_precision_gemm(a, w_q_packed, w_s, w_zp, size_bits, bias, group_size): positional order of the reference op
    params = list(inspect.signature(mixed_precision_gemm).parameters)
    assert params[:7] == ["x", "w_q_packed", "w_s", "w_zp", "weight_size_bits", "weight_bias", "group_size"]
    w = (10 * (torch.rand((256, 64)) - 0.3)).to(torch.float16)
    w_ref, w_q, w_s, w_zp = quantize_weights(w, scalar_types.uint4, 128, zero_points=True, ref_zero_points_after_scales=False)
    packed = pack_rows(w_q, 4, *w_q.shape)
    assert packed.shape == (32, 64) and packed.dtype == torch.int32 and w_s.shape == (2, 64) and w_zp.shape == (2, 64)
    # ref_zero_points_after_scales: q*s - zp*s, two roundings (reference conch/third_party/vllm/quant_utils.py:71-74)
    w_ref2, w_q2, w_s2, w_zp2 = quantize_weights(w, scalar_types.uint4, 128, zero_points=True, ref_zero_points_after_scales=True)
    assert torch.equal(w_q, w_q2) and torch.equal(w_s, w_s2) and torch.equal(w_zp, w_zp2)
    s = w_s2.repeat_interleave(128, dim=0)
    z = w_zp2.repeat_interleave(128, dim=0)
    assert torch.equal(w_ref2, w_q2.to(torch.float16) * s - z.to(torch.float16) * s)
    assert torch.equal(w_ref, (w_q - z).to(torch.float16) * s)
