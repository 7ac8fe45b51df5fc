"""Timing helpers shared by the sweep scripts (CUDA events, cold-L2 by rotating over distinct weight copies)."""

from __future__ import annotations

import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

L2_BYTES = 126 * 1024 * 1024


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def copies_for_cold_l2(nbytes: int, factor: float = 3.0, cap: int = 64) -> int:
    """How many distinct copies of a weight set are needed so that >= factor x L2 is touched between reuses."""
    return max(2, min(cap, int(factor * L2_BYTES / max(nbytes, 1)) + 1))


def time_rotating(fns, iters: int = 200, warmup: int = 20, graph: bool = True) -> float:
    """Average microseconds per call when cycling through `fns` (each bound to a distinct weight copy).
    With graph=True the whole rotation is captured once and replayed, so Python launch overhead is excluded."""
    n = len(fns)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for i in range(max(warmup, n)):
            fns[i % n]()
        stream.synchronize()
        reps = max(1, iters // n)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for f in fns:
                    f()
            g.replay()
            stream.synchronize()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(stream)
            for _ in range(reps):
                g.replay()
            stop.record(stream)
        else:
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(stream)
            for _ in range(reps):
                for f in fns:
                    f()
            stop.record(stream)
        stop.synchronize()
    return start.elapsed_time(stop) * 1e3 / (reps * n)
