"""scaled_gemm (W8A8) throughput: cb200_scaled_gemm (tcgen05 kind::i8 / kind::f8f6f4) on the reference test's largest shape and
on Llama-3-8B prefill shapes, next to torch._scaled_mm (cuBLASLt fp8) as the library baseline where torch offers it.

    python benchmarks/scaled_gemm_bench.py [--out gpurun_out/scaled_gemm_bench.jsonl]

CUDA-event timed, operands rotated over enough copies to exceed L2.  Roofline: 2MNK over the dense 8-bit tensor peak (nominal
4.5 PFLOP/s on B200: twice the measured bf16 burst peak is used as the measured stand-in) or bytes over the measured HBM copy
bandwidth, whichever is slower."""

from __future__ import annotations

import argparse
import json

import torch

from common import peaks
from conch_b200.ops.quantization.gemm import scaled_gemm

SHAPES = [(4096, 2048, 4096), (4096, 4096, 6144), (4096, 4096, 14336), (4096, 14336, 4096), (8192, 8192, 8192), (16, 4096, 14336)]


def timed(fn, n_sets, iters=20):
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        fn(it % n_sets)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/scaled_gemm_bench.jsonl")
    args = ap.parse_args()
    pk = peaks()
    peak8 = 2.0 * pk["tflops"]
    with open(args.out, "w") as f:
        for m, k, n in SHAPES:
            for kind in ("int8", "fp8"):
                nbytes = m * k + n * k + m * n * 2
                n_sets = max(2, min(8, int(3 * 126e6 / nbytes) + 1))
                sets = []
                for i in range(n_sets):
                    g = torch.Generator(device="cuda").manual_seed(i)
                    if kind == "int8":
                        a = torch.randint(-32, 32, (m, k), dtype=torch.int8, device="cuda", generator=g)
                        b = torch.randint(-32, 32, (n, k), dtype=torch.int8, device="cuda", generator=g).T
                    else:
                        a = (0.25 * torch.rand((m, k), device="cuda", generator=g)).to(torch.float8_e4m3fn)
                        b = (0.25 * torch.rand((n, k), device="cuda", generator=g)).to(torch.float8_e4m3fn).T
                    sets.append((a, b))
                sa = 0.25 * torch.rand((m, 1), device="cuda")
                sb = 0.25 * torch.rand((n, 1), device="cuda")
                us = timed(lambda i: scaled_gemm(sets[i][0], sets[i][1], sa, sb, torch.float16), n_sets)
                row = {"M": m, "K": k, "N": n, "kind": kind, "us": us, "tflops": 2.0 * m * n * k / us / 1e6}
                roof_us = max(2.0 * m * n * k / (peak8 * 1e6), nbytes / (pk["hbm_gbs"] * 1e3))
                row["roofline_us"] = roof_us
                row["frac"] = roof_us / us
                if kind == "fp8":
                    try:
                        us_lib = timed(lambda i: torch._scaled_mm(sets[i][0], sets[i][1], scale_a=sa, scale_b=sb.T, out_dtype=torch.float16), n_sets)
                        row["torch_scaled_mm_us"] = us_lib
                    except Exception as e:  # noqa: BLE001
                        row["torch_scaled_mm_us"] = None
                        row["torch_scaled_mm_error"] = str(e)[:80]
                f.write(json.dumps(row) + "\n")
                lib = f"   torch._scaled_mm {row['torch_scaled_mm_us']:8.1f} us" if row.get("torch_scaled_mm_us") else ""
                print(f"{m:5d} x {k:5d} x {n:5d} {kind:4s}: {us:8.1f} us  {row['tflops']:7.1f} TFLOP/s  {100 * row['frac']:5.1f}% of roofline{lib}", flush=True)


if __name__ == "__main__":
    main()
