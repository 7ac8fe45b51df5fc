"""CPU, world_size 2, gloo: the N-sharding + all-gather plumbing of conch_b200.sharded.

The per-rank GEMM is injected (the CPU oracle) because there is no GPU here; what is under test is the host logic:
column slicing of packed weights / scales / zero-points, shard-major gather layout, reassembly, scalar zero-points."""

from __future__ import annotations

import os
import pathlib
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _worker(rank: int, world: int, port: int, m: int, wname: str, zp: bool, scalar_zp: bool, ret) -> None:
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import mpgemm_oracle as oracle
    from conch_b200.sharded import ShardedMixedPrecisionLinear, shard_columns

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = oracle.make_problem(m, 256, 256, wname, 128, torch.float16, zp, seed=4)
        w_zp = p.w_zp
        if scalar_zp:
            w_zp = torch.tensor([[3]], dtype=torch.int32)
        full = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, w_zp, p.wtype.bits, p.wtype.bias, 128)
        layer = ShardedMixedPrecisionLinear(
            p.w_q_packed, p.w_s, w_zp, p.wtype.bits, p.wtype.bias, 128, gemm_fn=oracle.mixed_precision_gemm_ref
        )
        assert layer.w_q_packed.shape == (256 // p.wtype.pack_factor, 128) and layer.w_q_packed.is_contiguous()
        assert layer.w_s.shape == (2, 128)
        assert torch.equal(layer.w_q_packed, shard_columns(p.w_q_packed, rank, world))
        local = layer.local(p.x)
        assert torch.equal(local, full[:, rank * 128 : (rank + 1) * 128])  # no K split -> bit-identical
        out = layer(p.x)
        assert out.shape == (m, 256)
        assert torch.equal(out, full)
        out2 = layer(-p.x)  # the next call must not overwrite what the previous one returned (M = 1 used to be a view)
        assert torch.equal(out, full) and out2.data_ptr() != out.data_ptr()
        g = layer.gathered(p.x)
        assert g.shape == (world, m, 128) and torch.equal(g[1 - rank], full[:, (1 - rank) * 128 : (2 - rank) * 128])
        ga, work = layer.gathered_async(p.x)  # enqueue only; the view is valid after the wait
        assert work is not None
        work.wait()
        assert torch.equal(ga.permute(1, 0, 2).reshape(m, 256), full)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m", [1, 5])
@pytest.mark.parametrize("wname,zp,scalar_zp", [("uint4", True, False), ("uint8b128", False, False), ("uint4", False, True)])
def test_sharded_all_gather_world2(m, wname, zp, scalar_zp):
    port = 29500 + (os.getpid() + m * 7 + len(wname) + int(zp) * 3 + int(scalar_zp) * 11) % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, m, wname, zp, scalar_zp, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_columns_rejects_bad_splits():
    from conch_b200.sharded import shard_columns

    with pytest.raises(ValueError):
        shard_columns(torch.zeros((4, 192), dtype=torch.int32), 0, 2)  # 96 columns per rank: not a multiple of 64
    assert shard_columns(None, 0, 2) is None
