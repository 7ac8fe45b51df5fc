"""Timeline of the decode kernel (debug library: CB200_NVCC_EXTRA=-DCB200_TC_TRACE build, run with CB200_DEBUG_LIB=1).

Per CTA (warp 0): globaltimer at kernel entry, after the weight window + metadata copies were issued, after
griddepcontrol.wait, after the staged operands landed, after the first batch, at the end of the main loop, after the
split-K arrival.  Prints percentiles over CTAs relative to the earliest entry.
"""
import ctypes
import sys

import torch
import common  # noqa: F401
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

name = sys.argv[1] if len(sys.argv) > 1 else "gate_up"
k, n = {"gate_up": (4096, 28672), "down": (14336, 4096), "o": (4096, 4096), "qkv": (4096, 6144)}[name]
lib = _lib.load()
st = quant_utils.scalar_types.uint4
sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(4)]
x = torch.randn((1, k), device="cuda", dtype=torch.float16)
out = torch.empty((1, n), device="cuda", dtype=torch.float16)
nblk = 1 << 16
trace = torch.zeros((nblk, 8), dtype=torch.int64, device="cuda")
for s_ in sets:
    mixed_precision_gemm(x, s_[0], s_[1], s_[2], 4, 0, 128, out=out, variant=_lib.DECODE_MMA)
torch.cuda.synchronize()
lib.cb200_debug_set_decode_trace.argtypes = [ctypes.c_void_p]
lib.cb200_debug_set_decode_trace(trace.data_ptr())
s_ = sets[0]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
mixed_precision_gemm(x, s_[0], s_[1], s_[2], 4, 0, 128, out=out, variant=_lib.DECODE_MMA)
ev1.record()
torch.cuda.synchronize()
lib.cb200_debug_set_decode_trace(None)
t = trace.cpu()
used = t[:, 0] > 0
t = t[used]
t0 = int(t[:, 0].min())
print(f"{name}: {int(used.sum())} CTAs, event time {ev0.elapsed_time(ev1) * 1e3:.1f} us, SMs used {len(set(t[:, 7].tolist()))}")
names = ["entry", "win_issued", "meta_issued", "pdl_wait", "batch0", "loop_end", "staged"]
for c, nm in enumerate(names):
    v = (t[:, c][t[:, c] > 0] - t0).float()
    if len(v) == 0:
        continue
    q = torch.quantile(v, torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0]))
    print(f"  {nm:9s} min {q[0]:8.0f}  p10 {q[1]:8.0f}  p50 {q[2]:8.0f}  p90 {q[3]:8.0f}  max {q[4]:8.0f}  ns")
d = (t[:, 5] - t[:, 0]).float()
print(f"  per-CTA lifetime entry->loop_end: p50 {d.median():.0f} ns, max {d.max():.0f} ns")
d = (t[:, 5] - t[:, 4]).float()
print(f"  main loop after first batch:      p50 {d.median():.0f} ns")
# stragglers: do SMs that host more CTAs finish later?  (compute-bound main loop => yes)
sm = t[:, 7]
cnt = torch.bincount(sm, minlength=148)
loop = (t[:, 5] - t[:, 4]).float()
end = (t[:, 5] - t0).float()
for c in sorted(set(cnt.tolist())):
    sel = cnt[sm] == c
    if int(sel.sum()) == 0:
        continue
    print(f"  SMs hosting {c} CTAs: {int((cnt == c).sum()):3d} SMs, main loop p50 {loop[sel].median():7.0f} ns, loop_end p50 {end[sel].median():7.0f} max {end[sel].max():7.0f} ns")
