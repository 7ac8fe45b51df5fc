"""Time the UNMODIFIED reference (Conch's Triton `mixed_precision_gemm`, installed under baseline/_ref by
`pip install --no-deps --target baseline/_ref /root/reference`) next to conch_b200 on the same B200, same inputs, same
cold-L2 protocol -- the "kernel to beat" (BASELINE.md section 5).  Also checks both against a torch fp32 matmul.

    python benchmarks/reference_triton.py [--out gpurun_out/reference_triton.jsonl]
"""

from __future__ import annotations

import argparse
import json
import pathlib
import sys

import torch

from common import ROOT, copies_for_cold_l2, peaks, time_rotating

sys.path.insert(0, str(ROOT / "baseline" / "_ref"))

CASES = [
    ("8b_o", 1, 4096, 4096), ("8b_qkv", 1, 4096, 6144), ("8b_down", 1, 14336, 4096), ("8b_gate_up", 1, 4096, 28672),
    ("8b_gate_up", 16, 4096, 28672), ("8b_o", 64, 4096, 4096), ("8b_o", 256, 4096, 4096), ("8b_o", 1024, 4096, 4096),
    ("8b_ffn_prefill", 2048, 4096, 14336), ("8b_ffn_prefill", 4096, 4096, 14336), ("readme_h100_row", 4096, 8192, 4096),
]  # fmt: skip


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/reference_triton.jsonl")
    args = ap.parse_args()
    from conch.ops.quantization.gemm import mixed_precision_gemm as ref_gemm  # the reference, unmodified

    from conch_b200 import quant_utils
    from conch_b200.ops.quantization.gemm import dequantize_weights, mixed_precision_gemm

    pk = peaks()
    st = quant_utils.scalar_types.uint4
    with open(args.out, "w") as f:
        for name, m, k, n in CASES:
            wbytes = k * n // 2
            copies = min(8, copies_for_cold_l2(wbytes, factor=2.0)) if m > 16 else copies_for_cold_l2(wbytes)
            sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(copies)]
            x = torch.randn((m, k), device="cuda", dtype=torch.float16)
            out = torch.empty((m, n), device="cuda", dtype=torch.float16)
            s0 = sets[0]
            w = dequantize_weights(s0[0], s0[1], s0[2], 4, 0, 128)
            want = x.float() @ w.float()
            got_ref = ref_gemm(x, s0[0], s0[1], s0[2], 4, 0, 128)
            got_new = mixed_precision_gemm(x, s0[0], s0[1], s0[2], 4, 0, 128)
            err_ref = (got_ref.float() - want).abs().max().item()
            err_new = (got_new.float() - want).abs().max().item()
            ours = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out)) for s in sets]
            refs = [(lambda s=s: ref_gemm(x, s[0], s[1], s[2], 4, 0, 128)) for s in sets]
            iters = 200 if m <= 16 else max(2 * copies, 24)
            us_new = time_rotating(ours, iters=iters, warmup=copies)
            us_ref = time_rotating(refs, iters=iters, warmup=copies, graph=False)  # Triton launcher allocates: no graph
            us_new_eager = time_rotating(ours, iters=iters, warmup=copies, graph=False)
            flops, algo = 2.0 * m * n * k, wbytes + m * k * 2
            row = {
                "case": name, "M": m, "K": k, "N": n, "us_conch_b200": us_new, "us_conch_b200_eager": us_new_eager,
                "us_reference_triton": us_ref, "speedup_eager_vs_eager": us_ref / us_new_eager, "speedup": us_ref / us_new,
                "TFLOPS_conch_b200": flops / us_new / 1e6, "TFLOPS_reference": flops / us_ref / 1e6,
                "GBps_conch_b200": algo / us_new / 1e3, "GBps_reference": algo / us_ref / 1e3,
                "max_abs_err_conch_b200": err_new, "max_abs_err_reference": err_ref, "peaks": pk["source"],
            }  # fmt: skip
            f.write(json.dumps(row) + "\n")
            f.flush()
            print(f"{name:16s} M={m:5d} K={k:5d} N={n:5d}: conch_b200 {us_new:9.1f} us (eager {us_new_eager:9.1f}) | reference Triton {us_ref:9.1f} us | "
                  f"x{us_ref / us_new_eager:5.2f} eager-vs-eager | err {err_new:.3g} vs {err_ref:.3g}", flush=True)
            del sets
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
