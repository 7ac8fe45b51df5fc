"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

For every (activation dtype x weight type x zero-points x group_size) combination this runs the
reference's own quantiser, packer and PyTorch oracle on seeded CPU inputs:

    conch.third_party.vllm.quant_utils.quantize_weights   (quant_utils.py:14-101)
    conch.third_party.vllm.quant_utils.pack_rows          (quant_utils.py:104-125)
    torch.matmul(a, w_ref)                                (tests/mixed_precision_gemm_test.py:70)

and stores inputs + outputs as raw bit patterns (uint16 for fp16/bf16).  It also writes
`tests/golden/hashes.json`: sha256 digests of the reference's outputs at the reference test's own
shapes ((128,256,128) and (1024,1024,1024), tests/mixed_precision_gemm_test.py:36), which are too
large to commit as tensors.  `tests/test_oracle_golden.py` replays all of it through
oracle/mpgemm_oracle.py and requires bit-equality.
"""

from __future__ import annotations

import hashlib
import itertools
import json
import pathlib
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = pathlib.Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"

DTYPES = {"fp16": torch.float16, "bf16": torch.bfloat16}
WTYPES = ["uint4", "uint4b8", "uint8", "uint8b128"]


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def digest(t: torch.Tensor) -> str:
    t = t.contiguous()
    if t.dtype in (torch.float16, torch.bfloat16):
        raw = bits16(t).tobytes()
    else:
        raw = t.numpy().tobytes()
    return hashlib.sha256(raw).hexdigest()


def seeded_inputs(m: int, k: int, n: int, dtype: torch.dtype, seed: int, recipe: str):
    gen = torch.Generator().manual_seed(seed)
    if recipe == "reference":
        x = (10 * (torch.rand((m, k), generator=gen, dtype=torch.float32) - 0.3)).to(dtype)
        w = (10 * (torch.rand((k, n), generator=gen, dtype=torch.float32) - 0.3)).to(dtype)
    else:
        x = torch.randn((m, k), generator=gen, dtype=torch.float32).to(dtype)
        w = (0.02 * torch.randn((k, n), generator=gen, dtype=torch.float32)).to(dtype)
    return x, w


def main() -> None:
    sys.path.insert(0, REF)
    from conch.third_party.vllm.quant_utils import pack_rows, quantize_weights
    from conch.third_party.vllm.scalar_type import scalar_types

    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)

    # ---- small tensor fixtures -------------------------------------------------------
    m, k, n = 5, 256, 64
    index = []
    for dname, wname, zp, g, recipe in itertools.product(DTYPES, WTYPES, (True, False), (128, -1), ("reference",)):
        dtype = DTYPES[dname]
        wtype = getattr(scalar_types, wname)
        seed = len(index)
        x, w = seeded_inputs(m, k, n, dtype, seed, recipe)
        w_ref, w_q, w_s, w_zp = quantize_weights(w, wtype, g, zero_points=zp)
        packed = pack_rows(w_q, wtype.size_bits, k, n)
        out = torch.matmul(x, w_ref)  # the reference test's oracle expression
        name = f"{dname}_{wname}_{'zp' if zp else 'sym'}_g{g if g > 0 else 'K'}_{recipe}"
        arrays = {
            "x": bits16(x),
            "w": bits16(w),
            "w_q": w_q.numpy().astype(np.uint8),
            "w_q_packed": packed.numpy().astype(np.int32),
            "w_s": bits16(w_s),
            "w_ref": bits16(w_ref),
            "out_ref_matmul": bits16(out),
        }
        if w_zp is not None:
            arrays["w_zp"] = w_zp.numpy().astype(np.int32)
        np.savez_compressed(OUT / f"{name}.npz", **arrays)
        index.append(
            {
                "file": f"{name}.npz",
                "dtype": dname,
                "wtype": wname,
                "bits": wtype.size_bits,
                "bias": wtype.bias,
                "zero_points": zp,
                "group_size": g,
                "recipe": recipe,
                "seed": seed,
                "m": m,
                "k": k,
                "n": n,
            }
        )

    # ---- digests at the reference test's own shapes ------------------------------------
    hashes = []
    for (mm, kk, nn), dname, wname, zp in itertools.product(
        [(128, 256, 128), (1024, 1024, 1024)], DTYPES, WTYPES, (True, False)
    ):
        dtype = DTYPES[dname]
        wtype = getattr(scalar_types, wname)
        x, w = seeded_inputs(mm, kk, nn, dtype, 0, "reference")
        w_ref, w_q, w_s, w_zp = quantize_weights(w, wtype, 128, zero_points=zp)
        packed = pack_rows(w_q, wtype.size_bits, kk, nn)
        hashes.append(
            {
                "m": mm,
                "k": kk,
                "n": nn,
                "dtype": dname,
                "wtype": wname,
                "zero_points": zp,
                "group_size": 128,
                "seed": 0,
                "recipe": "reference",
                "sha256": {
                    "x": digest(x),
                    "w_q_packed": digest(packed),
                    "w_s": digest(w_s),
                    "w_zp": digest(w_zp) if w_zp is not None else None,
                    "w_ref": digest(w_ref),
                },
            }
        )

    (OUT / "index.json").write_text(json.dumps(index, indent=1) + "\n")
    (OUT / "hashes.json").write_text(json.dumps(hashes, indent=1) + "\n")
    total = sum(p.stat().st_size for p in OUT.iterdir())
    print(f"wrote {len(index)} fixtures + {len(hashes)} digests, {total / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__":
    main()
