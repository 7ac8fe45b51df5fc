"""CPU: pin the oracle (oracle/mpgemm_oracle.py) against outputs of the REAL reference.

tests/golden/*.npz and hashes.json were produced by oracle/make_golden.py, which imports the reference from
/root/reference (quantize_weights / pack_rows / torch.matmul(a, w_ref)).  Everything here is bit-exact.
"""

from __future__ import annotations

import hashlib

import numpy as np
import pytest
import torch

import mpgemm_oracle as oracle
from helpers import DTYPES, bits, golden_hashes, golden_id, golden_index, load_golden

INDEX = golden_index()


@pytest.mark.parametrize("entry", INDEX, ids=golden_id)
def test_quantiser_and_packer_match_reference(entry):
    g = load_golden(entry)
    wtype = oracle.WEIGHT_TYPES[entry["wtype"]]
    w_ref, w_q, w_s, w_zp = oracle.quantize_weights(g["w"], wtype, entry["group_size"], entry["zero_points"])
    assert torch.equal(bits(w_ref), bits(g["w_ref"]))
    assert torch.equal(w_q, g["w_q"])
    assert torch.equal(bits(w_s), bits(g["w_s"]))
    if entry["zero_points"]:
        assert torch.equal(w_zp, g["w_zp"])
    else:
        assert w_zp is None and g["w_zp"] is None
    packed = oracle.pack_rows(w_q, wtype.bits)
    assert packed.dtype == torch.int32
    assert torch.equal(packed, g["w_q_packed"])
    assert torch.equal(oracle.unpack_rows(packed, wtype.bits), g["w_q"])


@pytest.mark.parametrize("entry", INDEX, ids=golden_id)
def test_dequantize_is_bit_exact_vs_reference_w_ref(entry):
    g = load_golden(entry)
    w = oracle.dequantize(
        g["w_q_packed"], g["w_s"], g["w_zp"], entry["bits"], entry["bias"], entry["group_size"], DTYPES[entry["dtype"]]
    )
    assert torch.equal(bits(w), bits(g["w_ref"]))
    # zero-points as vLLM passes them (uint8) and in the activation dtype give the same bits
    if g["w_zp"] is not None:
        for zp in (g["w_zp"].to(torch.uint8), g["w_zp"].to(DTYPES[entry["dtype"]])):
            w2 = oracle.dequantize(g["w_q_packed"], g["w_s"], zp, entry["bits"], entry["bias"], entry["group_size"], DTYPES[entry["dtype"]])
            assert torch.equal(bits(w2), bits(g["w_ref"]))


@pytest.mark.parametrize("entry", INDEX, ids=golden_id)
def test_gemm_oracle_agrees_with_reference_test_expression(entry):
    """fp32-accumulated oracle vs the reference test's `torch.matmul(a, w_ref)` (computed by the reference run)."""
    g = load_golden(entry)
    out = oracle.mixed_precision_gemm_ref(
        g["x"], g["w_q_packed"], g["w_s"], g["w_zp"], entry["bits"], entry["bias"], entry["group_size"]
    )
    # the reference's own tolerance (tests/mixed_precision_gemm_test.py:76-77) ...
    atol = min(5e-2 * np.sqrt(entry["k"]), 1)
    torch.testing.assert_close(out.float(), g["out_ref_matmul"].float(), rtol=1e-1, atol=atol)
    # ... and the much tighter north_star one (half-precision matmul on CPU accumulates in fp32 too)
    torch.testing.assert_close(out.float(), g["out_ref_matmul"].float(), rtol=1e-2, atol=1e-2)
    # and the exact same expression evaluated here reproduces the stored bits
    again = oracle.reference_test_expression(g["x"], g["w_ref"])
    torch.testing.assert_close(again.float(), g["out_ref_matmul"].float(), rtol=1e-2, atol=1e-2)


def _digest(t: torch.Tensor) -> str:
    t = t.contiguous()
    if t.dtype in (torch.float16, torch.bfloat16):
        return hashlib.sha256(t.view(torch.int16).numpy().tobytes()).hexdigest()
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()


HASHES = golden_hashes()


@pytest.mark.parametrize(
    "h", HASHES, ids=lambda h: f"{h['m']}x{h['k']}x{h['n']}_{h['dtype']}_{h['wtype']}_{'zp' if h['zero_points'] else 'sym'}"
)
def test_reference_test_shapes_reproduce_reference_digests(h):
    """At the reference test's own shapes the oracle's tensors hash to what the reference produced."""
    p = oracle.make_problem(
        h["m"], h["k"], h["n"], h["wtype"], h["group_size"], DTYPES[h["dtype"]], h["zero_points"], seed=h["seed"],
        recipe=h["recipe"],
    )  # fmt: skip
    want = h["sha256"]
    assert _digest(p.x) == want["x"]
    assert _digest(p.w_q_packed) == want["w_q_packed"]
    assert _digest(p.w_s) == want["w_s"]
    assert (None if p.w_zp is None else _digest(p.w_zp)) == want["w_zp"]
    assert _digest(p.w_ref) == want["w_ref"]
    w = oracle.dequantize(p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, h["group_size"], DTYPES[h["dtype"]])
    assert _digest(w) == want["w_ref"]
