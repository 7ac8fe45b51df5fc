"""Per-layer table of the decode regime: every Llama-3-8B projection at M = 1 .. 32 on each kernel family that accepts it.

    python benchmarks/decode_table.py [--ms 1,4,8,16,32] [--out gpurun_out/decode_table.jsonl]

Cold-L2 protocol (common.time_rotating): every timed call reads a different copy of the weights (>= 3x L2 between reuses); the
rotation is captured in a CUDA graph with programmatic dependent launch, so a number is kernel time + back-to-back launch gap,
CUDA-event timed.  Roofline = max(bytes / measured HBM bandwidth, 2MNK / measured tensor peak).
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
FAMILIES = {"stream": _lib.DECODE_STREAM, "mma": _lib.DECODE_MMA, "tc_ts": _lib.TC_TS, "auto": _lib.AUTO}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/decode_table.jsonl")
    ap.add_argument("--shapes", default="o,qkv,down,gate_up")
    ap.add_argument("--ms", default="1,4,8,16,32")
    ap.add_argument("--families", default="stream,mma,tc_ts")
    ap.add_argument("--wtype", default="uint4")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--no-hint", action="store_true", help="stream kernel without the L2 evict_first policy on the weight tiles (A/B)")
    args = ap.parse_args()
    lib = _lib.load()
    pk = peaks()
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[args.dtype]
    st = getattr(quant_utils.scalar_types, args.wtype)
    lib.cb200_set_pdl(1)
    if args.no_hint:
        lib.cb200_set_stream_tuning(-224, 0)
    with open(args.out, "w") as f:
        for name, m in itertools.product(args.shapes.split(","), [int(v) for v in args.ms.split(",")]):
            k, n = SHAPES[name]
            wbytes = k * n * st.size_bits // 8
            copies = copies_for_cold_l2(wbytes)
            sets = [quant_utils.random_packed_weights(k, n, st, args.group, dtype, "cuda", seed=i) for i in range(copies)]
            x = torch.randn((m, k), device="cuda", dtype=dtype)
            out = torch.empty((m, n), device="cuda", dtype=dtype)
            algo = wbytes + m * k * 2
            roof_us = max(algo / (pk["hbm_gbs"] * 1e3), 2.0 * m * k * n / (pk["tflops"] * 1e6))
            cells = []
            for fam in args.families.split(","):
                v = FAMILIES[fam]
                if fam == "tc_ts" and m < 9:
                    continue
                try:
                    fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], st.size_bits, st.bias, args.group, out=out, variant=v)) for s in sets]
                    us = time_rotating(fns)
                except (ValueError, RuntimeError, NotImplementedError) as e:
                    print(f"skip {name} M={m} {fam}: {e}")
                    continue
                row = {"layer": name, "M": m, "K": k, "N": n, "family": fam, "us": us, "roofline_us": roof_us, "frac": roof_us / us,
                       "GBps": algo / us / 1e3, "wtype": args.wtype, "dtype": args.dtype, "group": args.group}
                f.write(json.dumps(row) + "\n")
                cells.append(f"{fam} {us:7.2f} us {100 * roof_us / us:5.1f}%")
            print(f"{name:8s} M={m:3d} ({algo / 1e6:5.1f} MB, roof {roof_us:6.2f} us):  " + "   ".join(cells), flush=True)
    lib.cb200_set_pdl(0)
    lib.cb200_set_stream_tuning(0, 0)


if __name__ == "__main__":
    main()
