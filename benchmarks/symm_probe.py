"""Does torch's symmetric memory (peer pointers + device-side barrier) work on this box?  torchrun --nproc-per-node 2."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import torch.distributed._symmetric_memory as symm_mem

t = symm_mem.empty((world, 4096), dtype=torch.float16, device=f"cuda:{rank}")
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "sig", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
t.zero_()
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (world, 4096), torch.float16)
peer[rank].fill_(float(rank + 1))   # P2P write into the peer's slot `rank`
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "got", t[:, 0].tolist(), flush=True)
# timing: peer copy + barrier, in a CUDA graph
y = torch.randn(4096, device=f"cuda:{rank}", dtype=torch.float16)
def step():
    peer[rank].copy_(y)
    hdl.barrier(channel=0)
for _ in range(5): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(50): step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(rank, f"graph: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us per (peer copy + barrier)", flush=True)
dist.barrier()
os._exit(0)
