"""Probe m_tile for the big prefill shapes (wave quantisation vs per-tile efficiency)."""
import torch
from common import time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

lib = _lib.load()
st = quant_utils.scalar_types.uint4
for name, m, k, n in (("8b_ffn", 4096, 4096, 14336), ("8b_ffn", 2048, 4096, 14336), ("readme", 4096, 8192, 4096), ("70b_shard8", 4096, 8192, 3584), ("8b_o", 4096, 4096, 4096), ("8b_o", 1024, 4096, 4096)):
    sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(6)]
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    out = torch.empty((m, n), device="cuda", dtype=torch.float16)
    for mt, cl in ((256, 1), (256, 2), (128, 1)):
        lib.cb200_set_tc_tuning(mt, 1)
        lib.cb200_set_tc_cluster(cl)
        fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=_lib.TC_TS)) for s in sets]
        us = time_rotating(fns, iters=24, warmup=6)
        lib.cb200_set_tc_tuning(0, 0)
        lib.cb200_set_tc_cluster(0)
        tiles = ((n + 127) // 128) * ((m + mt - 1) // mt)
        print(f"{name:10s} M={m} K={k} N={n} m_tile={mt:3d} cluster={cl}: {us:8.1f} us {2.0 * m * n * k / us / 1e6:7.1f} TFLOPS  tiles={tiles} waves={tiles / 148:.2f}", flush=True)
    del sets
    torch.cuda.empty_cache()
