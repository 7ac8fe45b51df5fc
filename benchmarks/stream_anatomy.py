"""Where does a streaming-decode launch spend its time?  Re-times one layer with parts of the kernel switched off
(cb200_debug_set_stream_switches: results are wrong, timings are what matters) and with different rings / grids.

    python benchmarks/stream_anatomy.py [--shape gate_up] [--m 1]
"""

from __future__ import annotations

import argparse

import torch

from common import copies_for_cold_l2, peaks, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm
from stream_sweep import SHAPES


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="gate_up")
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--pdl", type=int, default=1)
    args = ap.parse_args()
    lib = _lib.load()
    k, n = SHAPES[args.shape]
    st = quant_utils.scalar_types.uint4
    wbytes = k * n // 2
    sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(copies_for_cold_l2(wbytes))]
    x = torch.randn((args.m, k), device="cuda", dtype=torch.float16)
    out = torch.empty((args.m, n), device="cuda", dtype=torch.float16)
    fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=_lib.DECODE_STREAM)) for s in sets]
    pk = peaks()
    print(f"{args.shape} M={args.m}: {wbytes / 1e6:.1f} MB, roofline {wbytes / pk['hbm_gbs'] / 1e3:.2f} us, pdl={args.pdl}")
    lib.cb200_set_pdl(args.pdl)
    cases = [("full kernel", 0, 0, 0), ("consumers skip math", 1, 0, 0), ("no W loads", 2, 0, 0), ("no W loads, no math", 3, 0, 0),
             ("no x loads", 4, 0, 0), ("no meta loads", 8, 0, 0), ("W loads only, no math", 13, 0, 0), ("nothing loaded, no math", 15, 0, 0),
             ("full, ring 60 KB", 0, 60, 0), ("full, ring 200 KB", 0, 200, 0), ("full, 296 CTAs", 0, 0, 296), ("full, 74 CTAs", 0, 0, 74),
             ("W only no math, ring 200 KB", 13, 200, 0), ("W only no math, 296 CTAs", 13, 0, 296)]  # fmt: skip
    try:
        for label, dbg, ring, ctas in cases:
            lib.cb200_debug_set_stream_switches(dbg)
            lib.cb200_set_stream_tuning(ring, ctas)
            us = time_rotating(fns)
            print(f"  {label:32s} {us:7.2f} us   {wbytes / us / 1e3:7.1f} GB/s")
    finally:
        lib.cb200_debug_set_stream_switches(0)
        lib.cb200_set_stream_tuning(0, 0)
        lib.cb200_set_pdl(0)


if __name__ == "__main__":
    main()
