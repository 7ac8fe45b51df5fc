"""Small helpers the reference's tests and benchmarks import next to the op (conch/third_party/vllm/utils.py:21-25,
conch/utils/benchmark.py)."""

from __future__ import annotations

import random

import torch


def seed_everything(seed: int) -> None:
    """Seed Python's and torch's generators (reference conch/third_party/vllm/utils.py:21-25)."""
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def benchmark_it(fn, *, iters: int = 100, warmup: int = 10) -> dict:
    """Time `fn` with CUDA events after `warmup` untimed calls; returns min / median / mean / max in milliseconds.  The shape of
    the reference's BenchmarkMetadata-based helper (conch/utils/benchmark.py) reduced to what the op's benchmark prints."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return {"min_ms": times[0], "median_ms": times[len(times) // 2], "mean_ms": sum(times) / len(times), "max_ms": times[-1], "iters": iters}
