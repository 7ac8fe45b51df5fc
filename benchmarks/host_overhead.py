"""Host-side cost of one `mixed_precision_gemm` call (Python op + ctypes + C-ABI dispatch + cudaLaunchKernelEx).

A tiny problem keeps the GPU far ahead of the host, so wall time / calls is the launch path's CPU cost per call.
"""
import cProfile
import pstats
import time

import torch
import common  # noqa: F401
from conch_b200 import quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

st = quant_utils.scalar_types.uint4
wq, ws, wzp = quant_utils.random_packed_weights(256, 64, st, 128, torch.float16, "cuda", seed=0)
x = torch.randn((1, 256), device="cuda", dtype=torch.float16)
out = torch.empty((1, 64), device="cuda", dtype=torch.float16)
for _ in range(200):
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out)
torch.cuda.synchronize()
n = 20000
t0 = time.perf_counter()
for _ in range(n):
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host cost per call: {(t1 - t0) / n * 1e6:.2f} us (with out=);", flush=True)
t0 = time.perf_counter()
for _ in range(n):
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host cost per call: {(t1 - t0) / n * 1e6:.2f} us (allocating the output)", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(5000):
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
