#!/usr/bin/env python
"""How should an N-sharded decode token move its GEMV outputs between the GPUs?  (torchrun, N = 2 / 4 / 8)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        benchmarks/push_modes.py [--steps 20] [--modes local,fused,...] [--config5]

Every mode runs bench.py's step -- 32 layers x (qkv, o, gate_up, down) M=1 GEMVs, every rank an equally sized column shard of an
N-times wider layer -- from a CUDA graph with programmatic dependent launch and reports ms per token, max over ranks:

  local          the GEMVs alone (no gather, no barrier): the per-GPU floor
  local_bar      the GEMVs + the token's cross-GPU barrier (what the barrier alone costs)
  fused          cb200_mpgemm_push, AUTO kernel choice, + barrier                      (bench.py's path so far)
  fused_stream   the same with the streaming kernel forced
  fused_mma      the same with the register-window kernel forced
  fused_nobar    fused (AUTO) without the barrier (what the stores alone cost)
  async          GEMV stores locally, cb200_push per GEMV on a side stream with completion counters, + barrier
  async_nosig    the same without completion counters
  token          GEMVs store locally into one arena, ONE cb200_push of the whole token's outputs, + barrier

--config5 adds the 70B FFN shard chain at M=1 (K=8192, N=28672/G): wait kernel vs in-kernel wait (cb200_mpgemm_chain), and the
M=4096 gather: NCCL vs copy engines.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import BIAS, BITS, GROUP, LAYERS, PROJECTIONS  # noqa: E402
from conch_b200 import _lib, quant_utils  # noqa: E402
from conch_b200.ops.quantization.gemm import mixed_precision_gemm  # noqa: E402
from conch_b200.sharded import ShardedMixedPrecisionLinear  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--modes", default="local,local_bar,fused,fused_stream,fused_mma,fused_nobar,async,async_nosig,token")
    ap.add_argument("--config5", action="store_true")
    ap.add_argument("--layers", type=int, default=LAYERS)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.cb200_set_pdl(1)
    st = quant_utils.scalar_types.uint4
    nl = args.layers
    weights = [[quant_utils.random_packed_weights(k, n, st, GROUP, torch.float16, dev, seed=1000 * rank + 4 * layer + pi)
                for pi, (_, k, n) in enumerate(PROJECTIONS)] for layer in range(nl)]  # fmt: skip
    gen = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn((1, k), generator=gen, device=dev, dtype=torch.float16) for _, k, _ in PROJECTIONS]
    sharded = [[ShardedMixedPrecisionLinear(*weights[layer][pi], BITS, BIAS, GROUP, already_sharded=True) for pi in range(len(PROJECTIONS))]
               for layer in range(nl)]  # fmt: skip
    # one arena for the "token" mode: [G, total columns of the token] in symmetric memory
    import torch.distributed._symmetric_memory as symm_mem

    cols = [n for _, _, n in PROJECTIONS]
    tot = nl * sum(cols)
    arena = symm_mem.empty((world, tot), dtype=torch.float16, device=dev)
    ahdl = symm_mem.rendezvous(arena, dist.group.WORLD.group_name)
    arena_peers = [int(ahdl.buffer_ptrs[p]) + rank * tot * 2 for p in range(world) if p != rank]
    arena_outs, off = [], 0
    for layer in range(nl):
        row = []
        for n in cols:
            row.append(arena[rank, off : off + n].view(1, n))
            off += n
        arena_outs.append(row)
    outs = [[torch.empty((1, n), device=dev, dtype=torch.float16) for n in cols] for _ in range(nl)]

    def step(mode: str) -> None:
        for layer in range(nl):
            for pi in range(len(PROJECTIONS)):
                L = sharded[layer][pi]
                if mode in ("local", "local_bar"):
                    L.local(xs[pi], out=outs[layer][pi])
                elif mode in ("fused", "fused_nobar"):
                    L.pushed(xs[pi])
                elif mode == "fused_stream":
                    L.pushed(xs[pi], variant=_lib.DECODE_STREAM)
                elif mode == "fused_mma":
                    L.pushed(xs[pi], variant=_lib.DECODE_MMA)
                elif mode == "async":
                    L.pushed_async(xs[pi])
                elif mode == "async_nosig":
                    L.pushed_async(xs[pi], signalled=False)
                elif mode == "token":
                    L.local(xs[pi], out=arena_outs[layer][pi])
                else:
                    raise SystemExit(f"unknown mode {mode}")
        if mode == "token":
            peer_arr = (ctypes.c_void_p * len(arena_peers))(*arena_peers)
            _lib.raise_for_status(lib.cb200_push(arena[rank].data_ptr(), tot * 2, peer_arr, len(arena_peers), None, None,
                                                 torch.cuda.current_stream(dev).cuda_stream))  # fmt: skip
        if mode in ("async", "async_nosig"):
            ShardedMixedPrecisionLinear.join(dev)
        if mode not in ("local", "fused_nobar"):
            ahdl.barrier(channel=0)

    def timed_graph(fn, steps: int) -> float:
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            fn()
            fn()
            stream.synchronize()
            dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                fn()
            for _ in range(3):
                g.replay()
            stream.synchronize()
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                g.replay()
            e1.record(stream)
            e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t.item())

    res = {"world": world, "layers": nl}
    for mode in args.modes.split(","):
        if not mode:
            continue
        try:
            res[mode] = round(timed_graph(lambda: step(mode), args.steps), 4)
        except Exception as e:  # noqa: BLE001
            res[mode] = f"error: {type(e).__name__}: {e}"[:300]
        if rank == 0:
            print(json.dumps({mode: res[mode]}), flush=True)

    # correctness of the decoupled push and of the token arena: every rank's slot 0 rows equal that rank's local result
    if "async" in res and not isinstance(res["async"], str):
        L = sharded[0][2]
        buf = L.pushed_async(xs[2])
        ShardedMixedPrecisionLinear.join(dev)
        torch.cuda.synchronize(dev)
        dist.barrier()
        mine = buf[rank].clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ok = all(torch.equal(buf[r], gathered[r]) for r in range(world))
        res["async_values_match"] = bool(ok)

    if args.config5:
        res["config5"] = config5(world, rank, dev, lib)
    if rank == 0:
        print(json.dumps(res), flush=True)
    sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


def config5(world: int, rank: int, dev, lib) -> dict:
    k, n = 8192, 28672
    n_local = n // world
    st = quant_utils.scalar_types.uint4
    copies = max(4, min(24, int(2.2 * 126e6 / (k * n_local // 2)) + 1))
    copies += copies & 1
    layers = [ShardedMixedPrecisionLinear(*quant_utils.random_packed_weights(k, n_local, st, GROUP, torch.float16, dev, seed=900 + 31 * rank + i),
                                          BITS, BIAS, GROUP, already_sharded=True) for i in range(copies)]  # fmt: skip

    def timed(fn, iters: int, graph: bool) -> float:
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            for i in range(copies):
                fn(i)
            stream.synchronize()
            dist.barrier()
            g = None
            if graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for i in range(copies):
                        fn(i)
                g.replay()
                stream.synchronize()
            dist.barrier()
            torch.cuda.synchronize(dev)
            reps = max(1, iters // copies)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                if g is not None:
                    g.replay()
                else:
                    for i in range(copies):
                        fn(i)
            e1.record(stream)
            e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * copies)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t.item())

    out = {"copies": copies}
    x = torch.randn((1, k), device=dev, dtype=torch.float16)
    lib.cb200_set_pdl(1)
    louts = [torch.empty((1, n_local), device=dev, dtype=torch.float16) for _ in range(copies)]
    out["m1_compute_us"] = timed(lambda i: layers[i].local(x, out=louts[i]), 400, True)

    def fused_wait(i):
        layers[i].pushed_signalled(x)
        layers[i].wait()

    out["m1_push_signal_wait_kernel_us"] = timed(fused_wait, 400, True)

    def chained(i):
        # GEMV i's kernel waits, itself, for all ranks' slices of GEMV i-1 (per-GEMM visibility of the full [1, N] everywhere)
        layers[i].pushed_signalled(x, after=layers[i - 1])

    # prime the chain: layer copies-1 must have signalled once more than it has been waited for
    layers[copies - 1].pushed_signalled(x)
    torch.cuda.synchronize(dev)
    dist.barrier()
    try:
        out["m1_chain_us"] = timed(chained, 400, True)
    except Exception as e:  # noqa: BLE001
        out["m1_chain_us"] = f"error: {type(e).__name__}: {e}"[:300]
    # values: the chained result equals the local result
    torch.cuda.synchronize(dev)
    ref = layers[1].local(x)
    buf = layers[1]._symmetric(1, torch.float16, dev)[0]
    out["m1_chain_values_match"] = bool(torch.equal(buf[rank], ref))
    lib.cb200_set_pdl(0)

    m = 4096
    xb = torch.randn((m, k), device=dev, dtype=torch.float16)
    o1 = torch.empty((m, n_local), device=dev, dtype=torch.float16)
    out["m4096_compute_us"] = timed(lambda i: layers[i].local(xb, out=o1), 48, False)
    outs2 = [torch.empty((m, n_local), device=dev, dtype=torch.float16) for _ in range(2)]
    gbufs = [torch.empty((world * m, n_local), device=dev, dtype=torch.float16) for _ in range(2)]
    pending = [None, None]

    def piped(i):
        b = i & 1
        if pending[b] is not None:
            pending[b].wait()
        layers[i].local(xb, out=outs2[b])
        pending[b] = dist.all_gather_into_tensor(gbufs[b], outs2[b], async_op=True)

    out["m4096_nccl_piped_us"] = timed(piped, 48, False)
    for w in pending:
        if w is not None:
            w.wait()
    torch.cuda.synchronize(dev)
    try:  # the same on a process group whose NCCL stream has high priority
        from torch.distributed import ProcessGroupNCCL

        opts = ProcessGroupNCCL.Options(is_high_priority_stream=True)
        hp = dist.new_group(backend="nccl", pg_options=opts)
        pending = [None, None]

        def piped_hp(i):
            b = i & 1
            if pending[b] is not None:
                pending[b].wait()
            layers[i].local(xb, out=outs2[b])
            pending[b] = dist.all_gather_into_tensor(gbufs[b], outs2[b], group=hp, async_op=True)

        out["m4096_nccl_hiprio_piped_us"] = timed(piped_hp, 48, False)
        for w in pending:
            if w is not None:
                w.wait()
        torch.cuda.synchronize(dev)
    except Exception as e:  # noqa: BLE001
        out["m4096_nccl_hiprio_piped_us"] = f"error: {type(e).__name__}: {e}"[:300]
    if rank == 0:
        print(json.dumps({k_: v_ for k_, v_ in out.items() if k_.startswith("m4096")}), flush=True)
    for mode, ctas in (("ce", 0), ("p2p", 32), ("p2p", 64), ("multicast", 16), ("multicast", 32)):
        events = [None, None]

        def bulk(i, mode=mode, ctas=ctas):
            b = i & 1
            if events[b] is not None:
                torch.cuda.current_stream(dev).wait_event(events[b])  # slot b's previous gather is complete everywhere
            _, events[b] = layers[i].gathered_push_async(xb, slot=b, mode=mode, max_ctas=ctas)

        tag = f"m4096_{mode}{ctas or ''}_piped_us"
        try:
            out[tag] = timed(bulk, 48, False)
            for e in events:
                if e is not None:
                    torch.cuda.current_stream(dev).wait_event(e)
            torch.cuda.synchronize(dev)
            # values: slot r of the last-used buffer on this rank equals rank r's local result
            i_last = copies - 1
            buf = layers[i_last]._symm[(("bulk", i_last & 1, m), torch.float16)][0]
            mine = layers[i_last].local(xb)
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            out[tag.replace("_piped_us", "_values_match")] = bool(all(torch.equal(buf[r], gathered[r]) for r in range(world)))
        except Exception as e:  # noqa: BLE001
            out[tag] = f"error: {type(e).__name__}: {e}"[:300]
        if rank == 0:
            print(json.dumps({tag: out[tag]}), flush=True)
    return out


if __name__ == "__main__":
    main()
