"""Tiny driver for ncu: run one mixed_precision_gemm shape a few times (cold L2 via rotating weight copies).

    ncu --set full -k regex:decode_kernel -s 4 -c 2 -o gpurun_out/prof python benchmarks/profile_target.py --shape gate_up --m 1
"""

from __future__ import annotations

import argparse

import torch

from common import copies_for_cold_l2
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

SHAPES = {"o": (4096, 4096), "qkv": (4096, 6144), "down": (14336, 4096), "gate_up": (4096, 28672), "70b_ffn": (8192, 28672),
          "ffn_prefill": (4096, 14336), "readme": (8192, 4096)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="gate_up")
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--wtype", default="uint4")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    k, n = SHAPES[args.shape]
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    st = getattr(quant_utils.scalar_types, args.wtype)
    copies = min(4, copies_for_cold_l2(k * n * st.size_bits // 8))
    sets = [quant_utils.random_packed_weights(k, n, st, 128, dtype, "cuda", seed=i) for i in range(copies)]
    x = torch.randn((args.m, k), device="cuda", dtype=dtype)
    out = torch.empty((args.m, n), device="cuda", dtype=dtype)
    for i in range(args.iters):
        s = sets[i % copies]
        mixed_precision_gemm(x, s[0], s[1], s[2], st.size_bits, st.bias, 128, out=out, variant=args.variant)
    torch.cuda.synchronize()
    print("variant", _lib.VARIANT_NAMES[_lib.load().cb200_last_variant()])


if __name__ == "__main__":
    main()
