"""Sweep the decode kernel's tiling knobs on the Llama-3 decode shapes (BASELINE configs 2 and 4).

    python benchmarks/decode_sweep.py [--quick] [--out gpurun_out/decode_sweep.jsonl]

Cold-L2 protocol: every timed call reads a different copy of the weights (>= 3x L2 between reuses); the rotation is
captured in a CUDA graph so the number is kernel time + back-to-back launch gaps, CUDA-event timed.
"""

from __future__ import annotations

import argparse
import itertools
import json

import torch

from common import copies_for_cold_l2, peaks, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

SHAPES = {"o": (4096, 4096), "qkv": (4096, 6144), "down": (14336, 4096), "gate_up": (4096, 28672), "70b_ffn": (8192, 28672)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/decode_sweep.jsonl")
    ap.add_argument("--shapes", default="o,qkv,down,gate_up")
    ap.add_argument("--ms", default="1,4,8")
    ap.add_argument("--wtype", default="uint4")
    ap.add_argument("--dtype", default="fp16")
    args = ap.parse_args()
    lib = _lib.load()
    pk = peaks()
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    st = getattr(quant_utils.scalar_types, args.wtype)
    tilings = [(0, 0, 0, 0)]
    if not args.quick:
        tilings += [(mode, 0, upw, 0) for mode in (1, 2) for upw in (0, 1, 2, 3, 4, 8)]
    rows = []
    with open(args.out, "w") as f:
        for name, m in itertools.product(args.shapes.split(","), [int(v) for v in args.ms.split(",")]):
            k, n = SHAPES[name]
            wbytes = k * n * st.size_bits // 8
            copies = copies_for_cold_l2(wbytes)
            sets = [quant_utils.random_packed_weights(k, n, st, 128, dtype, "cuda", seed=i) for i in range(copies)]
            x = torch.randn((m, k), device="cuda", dtype=dtype)
            out = torch.empty((m, n), device="cuda", dtype=dtype)
            algo = wbytes + m * k * 2
            best = None
            for wn, wk, ks, vec in tilings:
                lib.cb200_set_decode_tuning(wn, wk, ks, vec)
                try:
                    fns = [
                        (lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], st.size_bits, st.bias, 128, out=out, variant=_lib.DECODE_MMA))
                        for s in sets
                    ]
                    us = time_rotating(fns)
                except (ValueError, RuntimeError) as e:
                    print(f"skip {name} M={m} tiling={(wn, wk, ks, vec)}: {e}")
                    continue
                finally:
                    lib.cb200_set_decode_tuning(0, 0, 0, 0)
                row = {
                    "shape": name, "M": m, "K": k, "N": n, "wtype": args.wtype, "dtype": args.dtype,
                    "warps_n": wn, "warps_k": wk, "split_k": ks, "vec": vec, "us": us,
                    "GBps": algo / us / 1e3, "frac_hbm": algo / us / 1e3 / pk["hbm_gbs"], "copies": copies,
                }  # fmt: skip
                f.write(json.dumps(row) + "\n")
                f.flush()
                if best is None or us < best["us"]:
                    best = row
            rows.append(best)
            print(f"BEST {name:8s} M={m:2d}: {best['us']:7.2f} us  {best['GBps']:7.1f} GB/s  {100 * best['frac_hbm']:5.1f}% of {pk['source']} HBM  "
                  f"tiling(wn,wk,ks,vec)=({best['warps_n']},{best['warps_k']},{best['split_k']},{best['vec']})", flush=True)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
