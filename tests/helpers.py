"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""

from __future__ import annotations

import json
import pathlib

import numpy as np
import torch

import mpgemm_oracle as oracle

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
DTYPES = {"fp16": torch.float16, "bf16": torch.bfloat16}

# north_star tolerance for the GEMM output against an fp32-accumulated reference
RTOL = 1e-2
ATOL = 1e-2


def from_bits(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(dtype)


def bits(t: torch.Tensor) -> torch.Tensor:
    return t.detach().cpu().contiguous().view(torch.int16)


def golden_index() -> list[dict]:
    return json.loads((GOLDEN / "index.json").read_text())


def golden_hashes() -> list[dict]:
    return json.loads((GOLDEN / "hashes.json").read_text())


def load_golden(entry: dict) -> dict:
    z = np.load(GOLDEN / entry["file"])
    dt = DTYPES[entry["dtype"]]
    out = {
        "x": from_bits(z["x"], dt),
        "w": from_bits(z["w"], dt),
        "w_q": torch.from_numpy(z["w_q"].astype(np.int32)),
        "w_q_packed": torch.from_numpy(z["w_q_packed"]),
        "w_s": from_bits(z["w_s"], dt),
        "w_ref": from_bits(z["w_ref"], dt),
        "out_ref_matmul": from_bits(z["out_ref_matmul"], dt),
        "w_zp": torch.from_numpy(z["w_zp"]) if "w_zp" in z.files else None,
    }
    return out


def golden_id(entry: dict) -> str:
    return entry["file"].removesuffix(".npz")


def to_cuda(*ts):
    return [None if t is None else t.cuda() for t in ts]


def assert_gemm_close(actual: torch.Tensor, expected: torch.Tensor, what: str = "") -> None:
    torch.testing.assert_close(
        actual.detach().cpu().float(), expected.detach().cpu().float(), rtol=RTOL, atol=ATOL, msg=lambda m: f"{what}: {m}"
    )


def make_problem(*args, **kwargs) -> oracle.Problem:
    return oracle.make_problem(*args, **kwargs)
