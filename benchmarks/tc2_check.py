"""Persistent tcgen05 variant (TC2_TS): correctness against TC_TS (and the fp32 torch reference) + timing."""
import torch
from common import time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import dequantize_weights, mixed_precision_gemm

st = quant_utils.scalar_types.uint4
for (m, k, n, g) in ((128, 512, 256, 128), (200, 1024, 320, 128), (1000, 2048, 1024, 64), (4096, 4096, 4096, 128), (4096, 4096, 4096, -1)):
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, g if g > 0 else k, torch.float16, "cuda", seed=m)
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    a = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, g, variant=_lib.TC_TS)
    b = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, g, variant=_lib.TC2_TS)
    b2 = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, g, variant=_lib.TC2_TS)
    torch.cuda.synchronize()
    w = dequantize_weights(wq, ws, wzp, 4, 0, g)
    ref = (x.float() @ w.float())
    print(f"M={m} K={k} N={n} g={g}: max|tc2-ts|={float((a.float() - b.float()).abs().max()):.4g} max|tc2-ref|={float((b.float() - ref).abs().max()):.4g} "
          f"max|ts-ref|={float((a.float() - ref).abs().max()):.4g} repeat_equal={bool(torch.equal(b, b2))}", flush=True)
for name, m, k, n in (("8b_ffn", 4096, 4096, 14336), ("8b_ffn", 2048, 4096, 14336), ("readme", 4096, 8192, 4096), ("8b_o", 4096, 4096, 4096), ("8b_o", 1024, 4096, 4096), ("8b_o", 512, 4096, 4096)):
    sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(6)]
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    out = torch.empty((m, n), device="cuda", dtype=torch.float16)
    for vname, v in (("tc_ts", _lib.TC_TS), ("tc2_ts", _lib.TC2_TS)):
        fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=v)) for s in sets]
        us = time_rotating(fns, iters=24, warmup=6)
        print(f"{name:8s} M={m} K={k} N={n} {vname:7s}: {us:8.1f} us {2.0 * m * n * k / us / 1e6:7.1f} TFLOPS", flush=True)
    del sets
    torch.cuda.empty_cache()
