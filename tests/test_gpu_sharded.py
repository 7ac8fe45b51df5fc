"""GPU, >= 2 B200s: N-sharded op + NCCL all-gather equals the single-GPU op bit for bit (skipped on a 1-GPU box)."""

from __future__ import annotations

import os
import pathlib
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, ret) -> None:
    import torch.distributed as dist

    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import mpgemm_oracle as oracle
    from conch_b200.ops.quantization.gemm import mixed_precision_gemm
    from conch_b200.sharded import ShardedMixedPrecisionLinear

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        for m in (1, 4, 16, 200):
            p = oracle.make_problem(m, 1024, 512, "uint4", 128, torch.float16, True, seed=m)
            x, wq, ws, wzp = (t.cuda() for t in (p.x, p.w_q_packed, p.w_s, p.w_zp))
            single = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
            layer = ShardedMixedPrecisionLinear(wq, ws, wzp, 4, 0, 128)
            out = layer(x)
            torch.cuda.synchronize()
            assert out.shape == single.shape
            assert torch.equal(out, single), f"sharded != single at M={m}"
            want = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, 4, 0, 128)
            torch.testing.assert_close(out.cpu().float(), want.float(), rtol=1e-2, atol=1e-2)
            if m <= ShardedMixedPrecisionLinear.PUSH_MAX_ROWS:
                # fused gather: the decode kernel's epilogue stores into every rank's buffer (cb200_mpgemm_push), one barrier
                buf = layer.pushed(x)
                layer.sync()
                torch.cuda.synchronize()
                fused = buf.permute(1, 0, 2).reshape(m, -1)
                assert torch.equal(fused, single), f"pushed != single at M={m}"
                # a second round through the same symmetric buffer, different activations
                layer.sync()  # nobody still reads round one
                buf = layer.pushed(-x)
                layer.sync()
                torch.cuda.synchronize()
                assert torch.equal(buf.permute(1, 0, 2).reshape(m, -1), mixed_precision_gemm(-x, wq, ws, wzp, 4, 0, 128))
                # per-GEMM visibility without a barrier: the kernel bumps a completion counter on every GPU, wait() spins on them
                layer.sync()
                for sign in (1.0, -1.0, 1.0):
                    buf = layer.pushed_signalled(sign * x)
                    layer.wait()
                    got = buf.permute(1, 0, 2).reshape(m, -1).clone()
                    layer.sync()  # (the test re-uses the buffer right away: nobody may still be reading it)
                    torch.cuda.synchronize()
                    assert torch.equal(got, mixed_precision_gemm(sign * x, wq, ws, wzp, 4, 0, 128)), f"pushed_signalled != single at M={m}"
                # decoupled push: the GEMV stores locally, cb200_push on a side stream carries the slice and raises the counters
                for sign in (1.0, -1.0):
                    buf = layer.pushed_async(sign * x)
                    layer.wait()
                    got = buf.permute(1, 0, 2).reshape(m, -1).clone()
                    ShardedMixedPrecisionLinear.join(x.device)
                    layer.sync()
                    torch.cuda.synchronize()
                    assert torch.equal(got, mixed_precision_gemm(sign * x, wq, ws, wzp, 4, 0, 128)), f"pushed_async != single at M={m}"
                # chain: the SECOND layer's kernel itself waits for the first layer's counters (cb200_mpgemm_chain): its activations
                # are the first layer's gathered output
                layer2 = ShardedMixedPrecisionLinear(wq[: 512 // 8], ws[: 512 // 128], wzp[: 512 // 128], 4, 0, 128)  # K = 512 = layer's N
                for sign in (1.0, -1.0):
                    buf = layer.pushed_signalled(sign * x)
                    hx = buf.view(1, -1) if m == 1 else None
                    if hx is not None:  # (M = 1: [G, 1, N/G] IS [1, N] in memory -- the gather buffer is the next layer's x)
                        buf2 = layer2.pushed_signalled(hx, after=layer)
                        layer2.wait()
                        got2 = buf2.permute(1, 0, 2).reshape(m, -1).clone()
                        layer.sync()
                        torch.cuda.synchronize()
                        want1 = mixed_precision_gemm(sign * x, wq, ws, wzp, 4, 0, 128)
                        want2 = mixed_precision_gemm(want1, wq[: 512 // 8].contiguous(), ws[: 512 // 128].contiguous(), wzp[: 512 // 128].contiguous(), 4, 0, 128)
                        assert torch.equal(got2, want2), f"chained != single at M={m}"
                    else:
                        layer.wait()
                        layer.sync()
                        torch.cuda.synchronize()
            else:
                # large M: the slice travels from a high-priority side stream (multicast stores / P2P stores / copy engines)
                for mode in ("auto", "p2p", "ce"):
                    buf, ev = layer.gathered_push_async(x, slot=0, mode=mode)
                    torch.cuda.current_stream().wait_event(ev)
                    torch.cuda.synchronize()
                    assert torch.equal(buf.permute(1, 0, 2).reshape(m, -1), single), f"gathered_push_async({mode}) != single at M={m}"
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_matches_single_gpu_nccl():
    import torch.multiprocessing as mp

    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29731, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
