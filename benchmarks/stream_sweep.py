"""Per-layer table of the decode regime: streaming kernel (CB200_DECODE_STREAM) against round 1's register-window kernel
(CB200_DECODE_MMA) on the Llama-3 decode shapes, M = 1..16, with and without programmatic dependent launch.

    python benchmarks/stream_sweep.py [--ms 1,4,8,16] [--shapes o,qkv,down,gate_up] [--rings 0] [--out gpurun_out/stream_sweep.jsonl]

Cold-L2 protocol: every timed call reads a different copy of the weights (>= 3x L2 between reuses); the rotation is captured
in a CUDA graph (PDL edges included), CUDA-event timed: the number is the steady-state time per GEMV inside a chain of GEMVs,
which is what a decode token is.
"""

from __future__ import annotations

import argparse
import itertools
import json

import torch

from common import copies_for_cold_l2, peaks, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

SHAPES = {"o": (4096, 4096), "qkv": (4096, 6144), "down": (14336, 4096), "gate_up": (4096, 28672), "70b_ffn": (8192, 28672),
          "70b_ffn_shard8": (8192, 3584)}  # fmt: skip


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/stream_sweep.jsonl")
    ap.add_argument("--shapes", default="o,qkv,down,gate_up")
    ap.add_argument("--ms", default="1,4,8,16")
    ap.add_argument("--wtype", default="uint4")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--rings", default="0", help="comma list of ring budgets in KiB (0 = default)")
    ap.add_argument("--ctas", default="0", help="comma list of CTA counts (0 = one per SM)")
    ap.add_argument("--no-mma", action="store_true", help="skip the round-1 kernel")
    ap.add_argument("--no-stream", action="store_true", help="skip the mma.sync streaming kernel")
    args = ap.parse_args()
    lib = _lib.load()
    pk = peaks()
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    st = getattr(quant_utils.scalar_types, args.wtype)
    rings = [int(v) for v in args.rings.split(",")]
    ctas = [int(v) for v in args.ctas.split(",")]
    with open(args.out, "a") as f:
        for name, m in itertools.product(args.shapes.split(","), [int(v) for v in args.ms.split(",")]):
            k, n = SHAPES[name]
            wbytes = k * n * st.size_bits // 8
            copies = copies_for_cold_l2(wbytes)
            sets = [quant_utils.random_packed_weights(k, n, st, args.group, dtype, "cuda", seed=i) for i in range(copies)]
            x = torch.randn((m, k), device="cuda", dtype=dtype)
            out = torch.empty((m, n), device="cuda", dtype=dtype)
            algo = wbytes + m * k * 2
            configs = []
            if not args.no_mma and m <= 16:
                configs += [("mma", _lib.DECODE_MMA, 1, 0, 0)]
            if not args.no_stream and m <= 16:
                configs += [("stream", _lib.DECODE_STREAM, 1, 0, 0)]
            for rk, ct in itertools.product(rings, ctas):
                configs += [("tcs", _lib.DECODE_TCS, 1, rk, ct)]
            configs += [("tcs-nopdl", _lib.DECODE_TCS, 0, 0, 0)]
            if m > 16:
                configs += [("tc_ts", _lib.TC_TS, 0, 0, 0)]
            line = f"{name:8s} M={m:2d} ({algo / 1e6:5.1f} MB, roof {algo / pk['hbm_gbs'] / 1e3:5.2f} us):"
            for label, variant, pdl, rk, ct in configs:
                lib.cb200_set_pdl(pdl)
                lib.cb200_set_stream_tuning(rk, ct)
                try:
                    fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], st.size_bits, st.bias, args.group, out=out, variant=variant))
                           for s in sets]  # fmt: skip
                    us = time_rotating(fns)
                except (ValueError, RuntimeError, NotImplementedError) as e:
                    print(f"skip {name} M={m} {label}: {e}")
                    continue
                finally:
                    lib.cb200_set_pdl(0)
                    lib.cb200_set_stream_tuning(0, 0)
                row = {"shape": name, "M": m, "K": k, "N": n, "wtype": args.wtype, "dtype": args.dtype, "g": args.group, "kernel": label,
                       "ring_kb": rk, "ctas": ct, "pdl": pdl, "us": us, "GBps": algo / us / 1e3,
                       "frac_hbm": algo / us / 1e3 / pk["hbm_gbs"], "copies": copies}  # fmt: skip
                f.write(json.dumps(row) + "\n")
                f.flush()
                tag = label + (f"[r{rk},c{ct}]" if (rk or ct) else "")
                line += f"  {tag} {us:6.2f} us {100 * row['frac_hbm']:5.1f}%"
            print(line, flush=True)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
