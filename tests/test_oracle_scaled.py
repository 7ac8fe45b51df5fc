"""The scaled-GEMM / quantiser oracle (oracle/scaled_oracle.py) pinned to the REAL reference: fixtures under tests/golden/scaled_*
were produced by importing /root/reference (oracle/make_golden_scaled.py).  CPU only."""

from __future__ import annotations

import pathlib

import numpy as np
import pytest
import torch

import scaled_oracle as so

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
OD = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}


def _t16(arr, dtype):
    return torch.from_numpy(arr.view(np.int16).copy()).view(dtype)


def _load_gemm(path):
    z = np.load(path)
    parts = path.stem.split("_")  # scaled_gemm_<kind>_<od>_<sA|vA>_<sB|vB>_<bias|nobias>
    kind, od = parts[2], OD[parts[3]]
    m, k, n = (int(v) for v in z["shape"])
    dt = torch.int8 if kind == "int8" else so.FP8
    a = torch.from_numpy(z["a"].copy()).view(dt).reshape(m, k)
    b = torch.from_numpy(z["b_t"].copy()).view(dt).reshape(n, k).T
    bias = _t16(z["bias"], od) if z["bias"].size else None
    return a, b, torch.from_numpy(z["scale_a"]), torch.from_numpy(z["scale_b"]), bias, od, _t16(z["ref"], od).reshape(m, n)


GEMM_FILES = sorted(GOLDEN.glob("scaled_gemm_*.npz"))
QUANT_FILES = sorted(GOLDEN.glob("scaled_quant_*.npz"))


def test_fixtures_exist():
    assert len(GEMM_FILES) == 32 and len(QUANT_FILES) == 6


@pytest.mark.parametrize("path", GEMM_FILES, ids=lambda p: p.stem)
def test_scaled_gemm_oracle_equals_the_reference(path):
    a, b, sa, sb, bias, od, ref = _load_gemm(path)
    got = so.scaled_gemm_reference(a, b, sa, sb, od, bias)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    # the op's own arithmetic (scale product first, bias added to the rounded value) stays within the reference test's
    # tolerance of it (tests/scaled_gemm_test.py:78), and in fact within two units in the last place
    op = so.scaled_gemm_op(a, b, sa, sb, od, bias)
    torch.testing.assert_close(op, ref, rtol=1e-1, atol=1e-1)
    ulp = 2.0**-10 if od == torch.float16 else 2.0**-7
    assert bool(((op.float() - ref.float()).abs() <= 2 * ulp * ref.float().abs().clamp_min(2.0**-14)).all())


@pytest.mark.parametrize("path", QUANT_FILES, ids=lambda p: p.stem)
def test_quantiser_oracles_equal_the_reference(path):
    z = np.load(path)
    dt = OD[path.stem.split("_")[2]]
    x8 = torch.from_numpy(z["x8"]) if dt == torch.float32 else _t16(z["x8"], dt)
    xf = torch.from_numpy(z["xf"]) if dt == torch.float32 else _t16(z["xf"], dt)
    s = torch.from_numpy(z["scale"])
    assert np.array_equal(so.scaled_int8_quant(x8, s).view(torch.uint8).numpy(), z["q8"])
    assert np.array_equal(so.scaled_fp8_quant(xf, s).view(torch.uint8).numpy(), z["qf"])
