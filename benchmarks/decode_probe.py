"""Probe decode-kernel variants (dequant / fused, split counts, PDL) on the Llama-3-8B decode shapes."""
import torch
from common import copies_for_cold_l2, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

lib = _lib.load()
st = quant_utils.scalar_types.uint4
for name, k, n in (("qkv", 4096, 6144), ("gate_up", 4096, 28672), ("down", 14336, 4096), ("o", 4096, 4096)):
    wbytes = k * n // 2
    copies = copies_for_cold_l2(wbytes)
    sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(copies)]
    x = torch.randn((1, k), device="cuda", dtype=torch.float16)
    out = torch.empty((1, n), device="cuda", dtype=torch.float16)
    for mode in (2,):
        for split in (0, 2, 4, 8, 16):
            for pdl in (0, 1):
                lib.cb200_set_decode_tuning(mode, 0, split, 0)
                lib.cb200_set_pdl(pdl)
                fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=_lib.DECODE_MMA)) for s in sets]
                us = time_rotating(fns)
                print(f"{name:8s} mode={'fused  ' if mode == 2 else 'dequant'} split={split:2d} pdl={pdl}: {us:7.2f} us  {wbytes / us / 1e3:7.1f} GB/s", flush=True)
    lib.cb200_set_decode_tuning(0, 0, 0, 0)
    lib.cb200_set_pdl(0)
    del sets
    torch.cuda.empty_cache()
