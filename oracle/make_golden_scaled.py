"""Generate tests/golden/scaled_*.npz by running the REAL reference (imported from /root/reference) on seeded CPU inputs.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden_scaled.py

  conch.reference.quantization.scaled_gemm._scaled_gemm_pytorch_ref   (scaled_gemm.py:12-27)
  conch.reference.quantization.int8._scaled_int8_quant_pytorch_ref    (int8.py:12-18)
  conch.reference.quantization.fp8._scaled_fp8_quant_pytorch_ref      (fp8.py:12-18)

Inputs come from oracle/scaled_oracle.make_scaled_problem (the reference test's recipe); 8-bit tensors are stored as raw bytes,
16-bit floats as uint16 bit patterns.  tests/test_oracle_scaled.py replays everything through oracle/scaled_oracle.py and requires
bit-equality."""

from __future__ import annotations

import itertools
import pathlib
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = pathlib.Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, REF)

import scaled_oracle as so  # noqa: E402
from conch.reference.quantization.fp8 import _scaled_fp8_quant_pytorch_ref  # noqa: E402
from conch.reference.quantization.int8 import _scaled_int8_quant_pytorch_ref  # noqa: E402
from conch.reference.quantization.scaled_gemm import _scaled_gemm_pytorch_ref  # noqa: E402


def raw(t: torch.Tensor) -> np.ndarray:
    t = t.contiguous()
    if t.element_size() == 1:
        return t.view(torch.uint8).numpy().copy()
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    return t.numpy().copy()


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    n_files = 0
    for kind, od, sa, sb, ub in itertools.product(("int8", "fp8"), ("fp16", "bf16"), (True, False), (True, False), (True, False)):
        out_dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[od]
        m, k, n = 24, 192, 40
        a, b, scale_a, scale_b, bias = so.make_scaled_problem(m, k, n, kind, out_dtype, sa, sb, ub, seed=7)
        ref = _scaled_gemm_pytorch_ref(a, b, scale_a, scale_b, out_dtype, bias)
        name = f"scaled_gemm_{kind}_{od}_{'sA' if sa else 'vA'}_{'sB' if sb else 'vB'}_{'bias' if ub else 'nobias'}"
        np.savez_compressed(OUT / f"{name}.npz", a=raw(a), b_t=raw(b.T), scale_a=scale_a.numpy(), scale_b=scale_b.numpy(),
                            bias=raw(bias) if ub else np.zeros(0, np.uint16), ref=raw(ref), shape=np.array([m, k, n]))
        n_files += 1
    for dt, scale in itertools.product(("fp16", "bf16", "fp32"), (0.1, 2.1)):
        dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[dt]
        g = torch.Generator().manual_seed(3)
        x8 = (torch.rand((7, 67), generator=g, dtype=torch.float32) * 1000 - 300).to(dtype)   # int8 recipe (+ negatives, clamps)
        xf = (torch.rand((7, 67), generator=g, dtype=torch.float32) * 4 - 1).to(dtype)        # fp8 recipe (+ negatives)
        xf[0, :4] = torch.tensor([1000.0, -1000.0, 0.0, 448.0 * scale]).to(dtype)             # saturation edge
        s = torch.tensor([scale], dtype=torch.float32)
        q8 = _scaled_int8_quant_pytorch_ref(x8, s)
        qf = _scaled_fp8_quant_pytorch_ref(xf, s)
        np.savez_compressed(OUT / f"scaled_quant_{dt}_s{str(scale).replace('.', 'p')}.npz", x8=raw(x8), xf=raw(xf), scale=s.numpy(),
                            q8=raw(q8), qf=raw(qf))
        n_files += 1
    print(f"wrote {n_files} fixtures to {OUT}")


if __name__ == "__main__":
    main()
