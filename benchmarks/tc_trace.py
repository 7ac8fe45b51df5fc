"""Per-role clock64 timeline of one CTA of the tcgen05 kernel (needs a library built with CB200_NVCC_EXTRA=-DCB200_TC_TRACE).

Columns (cycles relative to the first stamp): W producer got slot | x producer got slot | MMA: start wait, x_full, a_ready,
issued | dequant lane 0 of quarter 0: loop top, meta issued, w_full, dequantised, a_free, A stored.
"""
import ctypes
import sys

import torch
import common  # noqa: F401  (puts the repository root on sys.path)
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k, n = 4096, 14336
mt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib = _lib.load()
raw = ctypes.CDLL(str(_lib.LIB_PATH)) if hasattr(_lib, "LIB_PATH") else lib
st = quant_utils.scalar_types.uint4
wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=0)
x = torch.randn((m, k), device="cuda", dtype=torch.float16)
out = torch.empty((m, n), device="cuda", dtype=torch.float16)
trace = torch.zeros((k // 64, 16), dtype=torch.int64, device="cuda")
lib.cb200_set_tc_tuning(mt, 1)
lib.cb200_set_tc_cluster(1)
for _ in range(3):
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, variant=_lib.TC_TS)
torch.cuda.synchronize()
raw.cb200_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
raw.cb200_debug_set_tc_trace(trace.data_ptr())
mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, variant=_lib.TC_TS)
torch.cuda.synchronize()
raw.cb200_debug_set_tc_trace(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
names = ["Wslot", "Xslot", "Mwait", "Mxfull", "Maready", "Missued", "Dtop", "Dmeta", "Dwfull", "Ddeq", "Dafree", "Dst", "Mtrace2"]
print("it   " + " ".join(f"{s:>8s}" for s in names))
for it in range(k // 64):
    print(f"{it:3d}  " + " ".join(f"{int(t[it, c]) - t0:8d}" if t[it, c] > 0 else "       -" for c in range(13)))
