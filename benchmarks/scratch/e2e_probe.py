"""Where does the multi-GPU end-to-end step spend its host time?  (torchrun, N >= 2)"""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from bench import BIAS, BITS, GROUP, LAYERS, PROJECTIONS
from conch_b200 import _lib, quant_utils
from conch_b200.sharded import ShardedMixedPrecisionLinear

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load(); lib.cb200_set_pdl(1)
st = quant_utils.scalar_types.uint4
W = [[quant_utils.random_packed_weights(k, n, st, GROUP, torch.float16, dev, seed=1000 * rank + 4 * l + pi) for pi, (_, k, n) in enumerate(PROJECTIONS)] for l in range(LAYERS)]
xs = [torch.randn((1, k), device=dev, dtype=torch.float16) for _, k, _ in PROJECTIONS]
S = [[ShardedMixedPrecisionLinear(*W[l][pi], BITS, BIAS, GROUP, already_sharded=True) for pi in range(4)] for l in range(LAYERS)]
outs = [[torch.empty((1, n), device=dev, dtype=torch.float16) for _, _, n in PROJECTIONS] for _ in range(LAYERS)]
from conch_b200.ops.quantization.gemm import mixed_precision_gemm
def gemvs_local():
    for l in range(LAYERS):
        for pi in range(4):
            mixed_precision_gemm(xs[pi], *W[l][pi], BITS, BIAS, GROUP, out=outs[l][pi])
def gemvs_push():
    for l in range(LAYERS):
        for pi in range(4):
            S[l][pi].pushed(xs[pi])
for _ in range(3): gemvs_local(); gemvs_push()
S[0][0].sync(); torch.cuda.synchronize(); dist.barrier()
views = [S[l][pi].pushed_local_view(1) for l in range(LAYERS) for pi in range(4)]
flat = torch.empty((1, sum(v.shape[1] for v in views)), device=dev, dtype=torch.float16)
host = torch.empty(flat.shape, dtype=torch.float16).pin_memory()
def timed(fn, n=20):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) / n
    return round(t_host * 1e3, 3), round(t_all * 1e3, 3)
res = {}
res["local gemvs (host ms, host+device ms)"] = timed(gemvs_local)
res["push gemvs"] = timed(gemvs_push)
res["push gemvs + sync"] = timed(lambda: (gemvs_push(), S[0][0].sync()))
res["push gemvs + sync + cat"] = timed(lambda: (gemvs_push(), S[0][0].sync(), torch.cat(views, dim=1, out=flat)))
res["push gemvs + sync + cat + d2h"] = timed(lambda: (gemvs_push(), S[0][0].sync(), torch.cat(views, dim=1, out=flat), host.copy_(flat, non_blocking=True)))
res["cat alone"] = timed(lambda: torch.cat(views, dim=1, out=flat))
res["sync alone"] = timed(lambda: S[0][0].sync())
if rank == 0:
    for k, v in res.items(): print(k, v, flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
