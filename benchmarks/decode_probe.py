"""Probe what bounds the decode kernel on the gate_up shape: disable pieces through the debug bitmask."""
import torch
from common import copies_for_cold_l2, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

lib = _lib.load()
k, n, m = 4096, 28672, 1
st = quant_utils.scalar_types.uint4
wbytes = k * n // 2
copies = copies_for_cold_l2(wbytes)
sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(copies)]
x = torch.randn((m, k), device="cuda", dtype=torch.float16)
out = torch.empty((m, n), device="cuda", dtype=torch.float16)
names = {1: "no-x", 2: "no-meta", 4: "no-mma", 8: "no-publish"}
for mode in (2, 1):
    for split in (0, 1, 2, 4, 8):
        for dbg in (0, 1, 2, 4, 8, 15):
            lib.cb200_set_decode_tuning(mode, 0, split, dbg)
            fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=_lib.DECODE_MMA)) for s in sets]
            us = time_rotating(fns)
            label = "+".join(v for b, v in names.items() if dbg & b) or "full"
            print(f"mode={'fused  ' if mode == 2 else 'dequant'} split={split} {label:28s}: {us:7.2f} us  {wbytes / us / 1e3:7.1f} GB/s", flush=True)
lib.cb200_set_decode_tuning(0, 0, 0, 0)
