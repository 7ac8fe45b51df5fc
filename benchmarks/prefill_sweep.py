"""Time the tcgen05 variants on the prefill / mid-M shapes (BASELINE configs 3 and 4, plus the reference README shape).

    python benchmarks/prefill_sweep.py [--out gpurun_out/prefill_sweep.jsonl] [--quick]

Timing: CUDA events around a CUDA-graph replay of back-to-back launches over rotating weight copies (cold L2 for the
weights; x and the output are re-used).  TFLOPS = 2*M*N*K / t; the roofline is max(t_tensor, t_hbm) at the measured peaks.
"""

from __future__ import annotations

import argparse
import json

import torch

from common import copies_for_cold_l2, peaks, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

CASES = [
    # (name, M, K, N, wtype, dtype)
    ("8b_ffn_prefill", 4096, 4096, 14336, "uint4", "fp16"),
    ("8b_ffn_prefill", 2048, 4096, 14336, "uint4", "fp16"),
    ("readme_h100_row", 4096, 8192, 4096, "uint4b8", "fp16"),
    ("70b_ffn_shard8", 4096, 8192, 3584, "uint4", "fp16"),
]
MID = [("8b_o_mid", m, 4096, 4096, "uint4", "bf16") for m in (32, 64, 128, 256, 512, 1024)] + [
    ("8b_o_w8", m, 4096, 4096, "uint8b128", "bf16") for m in (64, 256, 1024, 4096)
]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/prefill_sweep.jsonl")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--variants", default="2,3")
    args = ap.parse_args()
    pk = peaks()
    lib = _lib.load()
    variants = [int(v) for v in args.variants.split(",")]
    cases = CASES if args.quick else CASES + MID
    with open(args.out, "w") as f:
        for name, m, k, n, wname, dname in cases:
            dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[dname]
            st = getattr(quant_utils.scalar_types, wname)
            zp = wname in ("uint4", "uint8")
            wbytes = k * n * st.size_bits // 8
            copies = min(8, copies_for_cold_l2(wbytes, factor=2.0))
            sets = [quant_utils.random_packed_weights(k, n, st, 128, dtype, "cuda", zero_points=zp, seed=i) for i in range(copies)]
            x = torch.randn((m, k), device="cuda", dtype=dtype)
            out = torch.empty((m, n), device="cuda", dtype=dtype)
            flops = 2.0 * m * n * k
            algo = wbytes + m * k * 2
            t_roof_us = max(flops / (pk["tflops"] * 1e12), algo / (pk["hbm_gbs"] * 1e9)) * 1e6
            for variant in variants:
                for mt in ([0] if m > 256 else [0]):
                    lib.cb200_set_tc_tuning(mt, 0)
                    try:
                        fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], st.size_bits, st.bias, 128, out=out, variant=variant)) for s in sets]
                        us = time_rotating(fns, iters=max(2 * copies, 24), warmup=copies)
                    except (ValueError, RuntimeError, NotImplementedError) as e:
                        print(f"skip {name} M={m} variant={variant}: {e}")
                        continue
                    finally:
                        lib.cb200_set_tc_tuning(0, 0)
                    row = {
                        "case": name, "M": m, "K": k, "N": n, "wtype": wname, "dtype": dname,
                        "variant": _lib.VARIANT_NAMES[variant], "m_tile": mt, "us": us, "TFLOPS": flops / us / 1e6,
                        "GBps": algo / us / 1e3, "roofline_us": t_roof_us, "frac_roofline": t_roof_us / us,
                        "frac_tensor_burst": flops / us / 1e6 / pk["tflops"], "peaks": pk["source"],
                    }  # fmt: skip
                    f.write(json.dumps(row) + "\n")
                    f.flush()
                    print(f"{name:16s} M={m:5d} K={k:5d} N={n:5d} {wname:9s} {dname} {row['variant']:6s}: {us:9.1f} us "
                          f"{row['TFLOPS']:7.1f} TFLOPS  {100 * row['frac_roofline']:5.1f}% of roofline ({pk['source']})", flush=True)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
