#!/usr/bin/env python
"""bench.py -- headline benchmark of the mixed-precision GEMM hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): one Llama-3-8B W4A16 decode token, M=1 -- the four quantised projections
(qkv 4096->6144, o 4096->4096, gate_up 4096->28672, down 14336->4096; uint4 + int32 zero-points, group_size 128,
fp16 activations) of all 32 layers = 128 GEMV launches reading 3.49 GB of packed weights (>> the 126 MB L2, so every
step streams its weights from HBM).  A "step" is one such token.

  value     whole-job HBM GB/s over the algorithmic bytes (packed weights + activations), inputs resident in HBM,
            the 128 launches replayed from a CUDA graph, CUDA-event timed, max over ranks.
  e2e       the same metric through the public Python op, call by call, x in pinned host memory, H2D of the step's
            activations and D2H of the step's outputs inside the timed region (wall clock, synchronised both sides).
  roofline  the decode kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference
            the reference's PyTorch path (`torch.matmul(x, w_ref)`, tests/mixed_precision_gemm_test.py:70) restated by
            the oracle, on this box's host cores, on a bounded sample (one layer's four projections per step).
Multi-GPU (N > 1): the headline line is weak-scaled -- every rank owns an equally sized column (N) shard of an N-times wider
layer, runs the same per-GPU step, and the outputs of every GEMV are gathered on every rank (fused into the kernel's epilogue).
Next to it, `config5_70b_ffn` reports BASELINE config 5 -- the Llama-3-70B FFN (K=8192, N=28672) column-sharded over the N GPUs,
STRONG scaling, at M=1 (fused gather) and M=4096 (multicast push overlapped with the next GEMM): compute / gather / end to end.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LAYERS = 32
# (name, K, N) of Llama-3-8B's quantised projections
PROJECTIONS = (("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096))
BITS, BIAS, GROUP = 4, 0, 128
M = 1
WORKLOAD = "llama3-8b-w4a16-decode-M1-32layers(qkv,o,gate_up,down;uint4+zp_int32;g128;fp16)"
METRIC = "W4A16 GEMM HBM GB/s, Llama-3-8B decode shapes M=1 (algorithmic bytes: packed weights + activations)"


def algorithmic_bytes(m: int = M) -> int:
    """B_W + B_X = K*N*bits/8 + M*K*2 per GEMM."""
    return LAYERS * sum(k * n * BITS // 8 + m * k * 2 for _, k, n in PROJECTIONS)


def flops(m: int = M) -> int:
    return LAYERS * sum(2 * m * k * n for _, k, n in PROJECTIONS)


def measured_peaks() -> tuple[float, str, dict]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML; nvidia-smi -lms cannot resolve a millisecond-scale timed region)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nv = None

    _NAMES = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }  # fmt: skip

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self._NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's PyTorch path restated by the oracle (the one place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, layers_per_step: int = LAYERS, budget_s: float = 150.0) -> dict:
    """The reference's PyTorch path on the host cores: `torch.matmul(x, w_ref)` with w_ref pre-dequantised offline exactly as
    tests/mixed_precision_gemm_test.py:70 does, restated by the oracle.  One step = `layers_per_step` layers x 4 projections
    (the GPU arm's step when layers_per_step == 32).  Two distinct layers' weights (872 MB of fp16, far beyond any host cache)
    are cycled, so every GEMV streams its weights from DRAM like the GPU arm does."""
    import torch

    sys.path.insert(0, str(ROOT / "oracle"))
    import mpgemm_oracle as oracle

    torch.set_num_threads(os.cpu_count() or 1)  # explicit: under torchrun OMP_NUM_THREADS defaults to 1
    threads = torch.get_num_threads()
    gen = torch.Generator().manual_seed(0)
    layer_sets = []
    for _ in range(2):
        layer = []
        for name, k, n in PROJECTIONS:
            packed = torch.randint(-(1 << 31), (1 << 31) - 1, (k // 8, n), generator=gen, dtype=torch.int64).to(torch.int32)
            w_s = (0.004 + 0.004 * torch.rand((k // GROUP, n), generator=gen)).to(torch.float16)
            w_zp = torch.randint(5, 11, (k // GROUP, n), generator=gen, dtype=torch.int32)
            w_ref = oracle.dequantize(packed, w_s, w_zp, BITS, BIAS, GROUP, torch.float16)  # offline, like the reference test
            x = torch.randn((M, k), generator=gen).to(torch.float16)
            layer.append((x, w_ref))
        layer_sets.append(layer)
    layer_bytes = sum(k * n * BITS // 8 + M * k * 2 for _, k, n in PROJECTIONS)

    def one_layer(i):
        for x, w_ref in layer_sets[i & 1]:
            oracle.reference_test_expression(x, w_ref)  # torch.matmul(a, w_ref), tests/mixed_precision_gemm_test.py:70

    def one_step():
        for i in range(layers_per_step):
            one_layer(i)

    t0 = time.perf_counter()
    one_layer(0)
    one_layer(1)
    per_layer = (time.perf_counter() - t0) / 2
    # keep the whole run inside the time budget: fewer timed steps if the host is slow (reported in `steps`)
    est_step = per_layer * layers_per_step
    steps = max(1, min(steps, int(budget_s / max(est_step, 1e-6)) - warmup))
    warmup = max(0, min(warmup, int(0.25 * budget_s / max(est_step, 1e-6))))
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return {
        "value": layer_bytes * layers_per_step / dt / 1e9,
        "unit": "GB/s",
        "cores": threads,
        "host_cpus": os.cpu_count(),
        "kind": "port",
        "sample": f"{layers_per_step} layers x 4 M=1 projections per step (the GPU arm's step has {LAYERS}), fp16 torch.matmul(x, w_ref) with "
        "w_ref pre-dequantised offline exactly as the reference test does, two distinct layers' weights cycled (872 MB); "
        f"torch.set_num_threads({threads}), no core pinning",
        "ms_per_step": dt * 1e3,
        "steps": steps,
        "warmup": warmup,
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": res["value"],
        "unit": "GB/s",
        "n_gpus": args.gpus,
        "steps": res["steps"],
        "warmup": res["warmup"],
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "M": M, "device": "host CPU", "l2": "n/a (872 MB of fp16 weights cycled through the host caches)",
                   "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus")},
        "e2e": {"value": res["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    from conch_b200 import _lib, quant_utils
    from conch_b200.build import build
    from conch_b200.ops.quantization.gemm import mixed_precision_gemm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build()  # no-op when the in-tree .so matches the sources
    if world > 1:
        dist.barrier()
    lib = _lib.load()
    # Programmatic dependent launch: each GEMV prefetches its (static) weights while the previous one drains.
    # Legal here because no kernel in the step writes weights (include/conch_b200.h, cb200_set_pdl).
    lib.cb200_set_pdl(0 if args.no_pdl else 1)
    push_variant = int(os.environ.get("CB200_BENCH_PUSH_VARIANT", "0"))  # debug: kernel family of the fused-gather GEMVs (0 = AUTO)

    # ---- synthetic model: 32 layers x 4 projections, drawn directly in the packed domain ------------------------
    st = quant_utils.scalar_types.uint4
    weights = []  # [layer][proj] -> (packed, w_s, w_zp)
    for layer in range(LAYERS):
        row = []
        for pi, (_, k, n) in enumerate(PROJECTIONS):
            row.append(quant_utils.random_packed_weights(k, n, st, GROUP, torch.float16, dev, seed=1000 * rank + 4 * layer + pi))
        weights.append(row)
    gen = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn((M, k), generator=gen, device=dev, dtype=torch.float16) for _, k, _ in PROJECTIONS]
    outs = [[torch.empty((M, n), device=dev, dtype=torch.float16) for _, _, n in PROJECTIONS] for _ in range(LAYERS)]
    sharded = None
    if world > 1:
        # every rank holds an equally sized column shard of a world-times wider layer (weak scaling)
        from conch_b200.sharded import ShardedMixedPrecisionLinear

        sharded = [
            [ShardedMixedPrecisionLinear(*weights[layer][pi], BITS, BIAS, GROUP, already_sharded=True) for pi in range(len(PROJECTIONS))]
            for layer in range(LAYERS)
        ]

    launches_per_step = 0

    def step(count: bool = False):
        nonlocal launches_per_step
        pending = None  # the previous GEMM's all-gather: enqueued on NCCL's stream, waited for one launch later
        for layer in range(LAYERS):
            for pi in range(len(PROJECTIONS)):
                packed, w_s, w_zp = weights[layer][pi]
                if world > 1 and args.nccl_gather:
                    # GEMM on the shard + NCCL all-gather of its output; the transfer overlaps the next GEMM
                    _, work = sharded[layer][pi].gathered_async(xs[pi], out=outs[layer][pi])
                    if pending is not None:
                        pending.wait()
                    pending = work
                elif world > 1:
                    # fused gather: the GEMV kernel's epilogue stores its slice into every rank's gather buffer (P2P)
                    sharded[layer][pi].pushed(xs[pi], variant=push_variant)
                else:
                    mixed_precision_gemm(xs[pi], packed, w_s, w_zp, BITS, BIAS, GROUP, out=outs[layer][pi])
                if count:
                    launches_per_step += lib.cb200_last_launch_count()
        if pending is not None:
            pending.wait()
        if world > 1 and not args.nccl_gather:
            sharded[0][0].sync()  # ONE cross-GPU barrier per token: every rank now holds all 128 gathered outputs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then capture one step into a CUDA graph ---------------------------------------------
    step(count=True)
    decode_variant = lib.cb200_last_variant()  # the kernel family AUTO dispatched the step's GEMVs to
    for _ in range(max(0, args.warmup - 1)):
        step()
    barrier()
    stream = torch.cuda.Stream(device=dev)
    use_graph = not args.no_graph
    graph = None
    if use_graph:
        with torch.cuda.stream(stream):
            step()
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                step()
            for _ in range(max(3, args.warmup)):
                graph.replay()
        barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ------------------------------------
    sampler = ClockSampler(local_rank)  # (NVML initialised BEFORE the barrier: tens of milliseconds, different on every rank)
    with torch.cuda.stream(stream):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if world > 1 and not args.nccl_gather and use_graph:
            # device-side cross-GPU barrier in front of the start event: the ranks' HOST threads leave the barrier above up to
            # milliseconds apart, and a rank that enqueues early would count its wait for the last one at the first token's
            # barrier as step time (20 steps: +8 % at 8 GPUs).  The timed region starts when every GPU is there.
            sharded[0][0].sync()
        with sampler as clocks:
            start.record(stream)
            for _ in range(args.steps):
                if graph is not None:
                    graph.replay()
                else:
                    step()
            stop.record(stream)
            stop.synchronize()
        barrier()
    ms = start.elapsed_time(stop)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    bytes_step = algorithmic_bytes() * world
    value = bytes_step / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public op: pinned host activations in, outputs back, every step ----------------
    x_host = [x.cpu().pin_memory() for x in xs]
    out_dev = torch.empty((LAYERS, sum(n for _, _, n in PROJECTIONS)), device=dev, dtype=torch.float16)
    out_host = torch.empty(out_dev.shape, dtype=torch.float16).pin_memory()
    col0 = [0]
    for _, _, n in PROJECTIONS:
        col0.append(col0[-1] + n)
    out_views = [[out_dev[layer : layer + 1, col0[pi] : col0[pi + 1]].view(M, -1) for pi in range(len(PROJECTIONS))] for layer in range(LAYERS)]
    # column slices of a row-major buffer are not contiguous for M>1, but with M=1 each view is one dense row segment
    x_dev = [torch.empty_like(x) for x in xs]

    local_views = None  # (N > 1) this rank's slots of the 128 gather buffers: they exist after the steps above
    if world > 1 and not args.nccl_gather:
        local_views = [sharded[layer][pi].pushed_local_view(M) for layer in range(LAYERS) for pi in range(len(PROJECTIONS))]
    out_flat = out_dev.view(1, -1)

    def e2e_step():
        for pi in range(len(PROJECTIONS)):
            x_dev[pi].copy_(x_host[pi], non_blocking=True)
        pending = None
        for layer in range(LAYERS):
            for pi in range(len(PROJECTIONS)):
                packed, w_s, w_zp = weights[layer][pi]
                if world > 1 and args.nccl_gather:
                    _, work = sharded[layer][pi].gathered_async(x_dev[pi], out=out_views[layer][pi])
                    if pending is not None:
                        pending.wait()
                    pending = work
                elif world > 1:
                    sharded[layer][pi].pushed(x_dev[pi], variant=push_variant)
                else:
                    mixed_precision_gemm(x_dev[pi], packed, w_s, w_zp, BITS, BIAS, GROUP, out=out_views[layer][pi])
        if pending is not None:
            pending.wait()
        if world > 1 and not args.nccl_gather:
            sharded[0][0].sync()
            # this rank's own slices (the result it computed) go back to the host, like the single-GPU arm
            torch.cat(local_views, dim=1, out=out_flat)
        out_host.copy_(out_dev, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        e2e_step()
    # wall clock through Python, max over ranks; best of three passes of e2e_steps steps -- the host side shares 16 cores with the
    # other ranks' launch, NCCL and NVML threads, and a single 40 ms stall in one rank doubles a 20-step average
    e2e_passes = 3
    t = None
    for _ in range(e2e_passes):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        tp = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        t = tp if t is None or float(tp.item()) < float(t.item()) else t
    e2e_value = bytes_step / float(t.item()) / 1e9
    h2d = sum(x.numel() * 2 for x in xs)
    d2h = out_dev.numel() * 2

    # ---- BASELINE config 5 (70B FFN, N-sharded, strong scaling): every rank takes part ------------------------------
    peak, peak_src, peaks = measured_peaks()
    used_graph = graph is not None
    d2h = out_dev.numel() * 2
    cfg5 = None
    prefill = prefill_probe(dev, peaks) if (world == 1 and rank == 0) else None  # before the sustained config-5 load: a burst figure
    if not args.no_config5:
        # release the 8B model first (rebinding, not `del`: the step closures above capture these names)
        weights.clear()
        outs.clear()
        out_views.clear()
        if sharded is not None:
            sharded.clear()
        out_dev = None
        graph = None
        torch.cuda.empty_cache()
        cfg5 = config5_probe(world, rank, dev, peaks)

    # ---- report --------------------------------------------------------------------------------------------------
    if rank == 0:
        per_gpu = value / world
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("decode_kernel_dram_bytes_per_launch")
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "GB/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f16",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "M": M,
                "launch": ("cuda-graph replay of the 128 GEMV launches" if used_graph else "eager python launches")
                + ("" if args.no_pdl else ", programmatic dependent launch"),
                "l2": "inputs larger than L2: 3.49 GB of packed weights per step vs 126 MB L2",
                "sharding": "none" if world == 1 else (f"N-sharded x{world}, NCCL all_gather_into_tensor after every GEMM (async: the gather of GEMM i overlaps GEMM i+1)" if args.nccl_gather else f"N-sharded x{world}, fused gather: every GEMV kernel stores its slice into all {world} ranks' gather buffers (P2P over NVLink, cb200_mpgemm_push), one cross-GPU barrier per token"),
            },
            "tflops": flops() * world / (ms_per_step * 1e-3) / 1e12,
            "roofline": {
                "bound": "hbm",
                "achieved": per_gpu,
                "peak": peak,
                "unit": "GB/s",
                "frac": per_gpu / peak,
                "traffic": traffic,
                "peak_source": peak_src,
                "kernel": ("cb200::stream_kernel<__half,4,1>" if decode_variant == _lib.DECODE_STREAM else "cb200::decode_kernel<__half,4,1,4,CB200_ZP_I32,FUSED=true>") + " (" + _lib.VARIANT_NAMES[decode_variant] + ")",
                "algorithmic_bytes_per_launch": algorithmic_bytes() / (LAYERS * len(PROJECTIONS)),
                "avg_launch_us": ms_per_step * 1e3 / (LAYERS * len(PROJECTIONS)),
            },
            "e2e": {
                "value": e2e_value,
                "unit": "GB/s",
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "ms_per_step": float(t.item()) * 1e3,
                "steps": e2e_steps,
                "passes": e2e_passes,
                "path": "conch_b200.mixed_precision_gemm called 128x per step from Python, pinned-host x -> device, outputs -> pinned host",
            },
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks.summary(),
        }
        if cfg5 is not None:
            line["config5_70b_ffn"] = cfg5
        if prefill is not None:
            line["prefill"] = prefill
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_reference(steps=3, warmup=1, layers_per_step=8, budget_s=25.0)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cpus")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL collectives captured in the CUDA graph keep communicator resources alive; tearing the process group
        # down under them can hang.  Everything is measured and printed: drain, then leave without the teardown.
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def config5_probe(world: int, rank: int, dev, peaks: dict) -> dict:
    """BASELINE config 5: Llama-3-70B W4A16 FFN (K=8192, N=28672, uint4 + int32 zero-points, g128, fp16), the weight
    column-(N-)sharded over the `world` GPUs -- STRONG scaling: the layer is fixed, every rank owns N/world columns -- with an
    all-gather of the outputs so that every rank ends up with the full [M, N].  Reported per BASELINE.md section 3:
    compute-only (the rank's shard GEMM), gather-only, and end to end, CUDA-event timed, max over ranks, cold L2 (rotating
    weight copies > 2 x L2).
      M = 1    : the gather is FUSED into the GEMV (cb200_mpgemm_push_signal: the kernel's epilogue stores its slice into every
                 rank's buffer over NVLink and then raises a completion flag on every GPU); every GEMV is followed by a wait for
                 all ranks' flags, i.e. the time includes per-GEMM visibility of the full [1, N] on every rank;
                 gather-only = one NCCL all_gather_into_tensor of the same 57 KB, for comparison.
      M = 4096 : tcgen05 GEMM on the shard writing straight into a symmetric gather buffer + cb200_push_bulk of the slice from a
                 high-priority side stream (NVLS multicast stores), software-pipelined: the gather of GEMM i overlaps GEMM i+1
                 (an NCCL all-gather kernel does not: it runs in the next GEMM's tail, profiles/r02_push_modes_n8_gather.log)."""
    import torch
    import torch.distributed as dist

    from conch_b200 import _lib, quant_utils
    from conch_b200.sharded import ShardedMixedPrecisionLinear

    lib = _lib.load()
    k, n = 8192, 28672
    n_local = n // world
    st = quant_utils.scalar_types.uint4
    shard_bytes = k * n_local // 2
    copies = max(3, min(24, int(2.2 * 126e6 / shard_bytes) + 1))
    copies += copies & 1  # even: the M=4096 pipeline alternates two gather slots
    layers = [ShardedMixedPrecisionLinear(*quant_utils.random_packed_weights(k, n_local, st, GROUP, torch.float16, dev, seed=900 + 31 * rank + i),
                                          BITS, BIAS, GROUP, already_sharded=True) for i in range(copies)]  # fmt: skip
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tc = float(peaks.get("bf16_tflops", 1590.0))

    def timed(fn, iters: int, graph: bool) -> float:
        """average microseconds per call of fn(i), max over ranks"""
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            for i in range(copies):
                fn(i)
            stream.synchronize()
            g = None
            if graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for i in range(copies):
                        fn(i)
                g.replay()
                stream.synchronize()
            if world > 1:
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
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"workload": f"llama3-70b-w4a16-ffn K={k} N={n} sharded over {world} GPU(s) along N (strong scaling), uint4+zp_int32, g128, fp16",
           "weight_copies_rotated": copies}  # fmt: skip
    for m in (1, 4096):
        x = torch.randn((m, k), device=dev, dtype=torch.float16)
        algo_bytes = k * n // 2 + m * k * 2
        flop = 2.0 * m * k * n
        t_roof = max(flop / (world * tc * 1e12), (k * n // 2 + world * m * k * 2) / (world * hbm * 1e9)) * 1e6
        row = {"roofline_us": t_roof, "gathered_output_bytes": m * n * 2}
        if m == 1:
            lib.cb200_set_pdl(1)
            outs = [torch.empty((m, n_local), device=dev, dtype=torch.float16) for _ in range(copies)]
            row["compute_us"] = timed(lambda i: layers[i].local(x, out=outs[i]), 400, graph=True)
            if world > 1:
                def fused(i):
                    layers[i].pushed_signalled(x)  # the kernel raises this rank's completion flag on every GPU ...
                    layers[i].wait()               # ... and the stream continues once all ranks' flags of GEMV i are here

                row["e2e_us"] = timed(fused, 400, graph=True)
                gbuf = torch.empty((world * m, n_local), device=dev, dtype=torch.float16)
                row["gather_us"] = timed(lambda i: dist.all_gather_into_tensor(gbuf, outs[i]), 200, graph=False)
                row["gather"] = ("fused into the GEMV epilogue (P2P vector stores over NVLink); per-GEMV visibility through completion flags "
                                 "raised by the kernel on every GPU + a one-warp wait (no barrier, no NCCL); gather_us = a stand-alone NCCL "
                                 "all_gather of the same bytes, for comparison")
            else:
                row["e2e_us"] = row["compute_us"]
                row["gather_us"] = 0.0
            lib.cb200_set_pdl(0)
        else:
            out = torch.empty((m, n_local), device=dev, dtype=torch.float16)
            row["compute_us"] = timed(lambda i: layers[i].local(x, out=out), 48, graph=False)
            if world > 1:
                outs2 = [torch.empty((m, n_local), device=dev, dtype=torch.float16) for _ in range(2)]
                gbufs = [torch.empty((world * m, n_local), device=dev, dtype=torch.float16) for _ in range(2)]
                row["gather_us"] = timed(lambda i: dist.all_gather_into_tensor(gbufs[0], outs2[0]), 24, graph=False)
                events = [None, None]

                def piped(i):
                    b = i & 1
                    if events[b] is not None:
                        torch.cuda.current_stream(dev).wait_event(events[b])  # slot b's previous gather is complete on every rank
                    # GEMM into this rank's slot of the symmetric [G, M, N/G] buffer, then the slice goes to every peer from a
                    # high-priority side stream (multimem.st to the NVLS multicast address where there is one) + a device barrier
                    _, events[b] = layers[i].gathered_push_async(x, slot=b)

                row["e2e_us"] = timed(piped, 48, graph=False)
                for e in events:
                    if e is not None:
                        torch.cuda.current_stream(dev).wait_event(e)
                torch.cuda.synchronize(dev)
                row["gather"] = ("cb200_push_bulk on a high-priority side stream (NVLS multicast stores where available, else P2P stores), "
                                 "double-buffered: the gather of GEMM i overlaps GEMM i+1; gather_us = a stand-alone NCCL all_gather_into_tensor "
                                 "of the same bytes, for comparison")
            else:
                row["e2e_us"] = row["compute_us"]
                row["gather_us"] = 0.0
        row["tflops_e2e"] = flop / row["e2e_us"] / 1e6
        row["GBps_e2e"] = algo_bytes / row["e2e_us"] / 1e3
        row["roofline_frac_e2e"] = t_roof / row["e2e_us"]
        row["roofline_frac_compute"] = t_roof / row["compute_us"]
        res[f"M{m}"] = row
        del x
    del layers
    torch.cuda.empty_cache()
    return res


def prefill_probe(dev, peaks: dict) -> dict:
    """Second regime of the same op, reported next to the decode headline: M=4096 x (K=4096, N=14336) W4A16 through the
    tcgen05 kernel, CUDA-event timed over weight copies that rotate (6 x 29 MB of packed weights + 33 MB of x per call)."""
    import torch

    from conch_b200 import _lib, quant_utils
    from conch_b200.ops.quantization.gemm import mixed_precision_gemm

    m, k, n = 4096, 4096, 14336
    st = quant_utils.scalar_types.uint4
    sets = [quant_utils.random_packed_weights(k, n, st, GROUP, torch.float16, dev, seed=100 + i) for i in range(6)]
    x = torch.randn((m, k), device=dev, dtype=torch.float16)
    out = torch.empty((m, n), device=dev, dtype=torch.float16)
    for s_ in sets:
        mixed_precision_gemm(x, s_[0], s_[1], s_[2], BITS, BIAS, GROUP, out=out)
    torch.cuda.synchronize(dev)
    iters = 24
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(iters):
        s_ = sets[i % len(sets)]
        mixed_precision_gemm(x, s_[0], s_[1], s_[2], BITS, BIAS, GROUP, out=out)
    stop.record()
    stop.synchronize()
    us = start.elapsed_time(stop) * 1e3 / iters
    tflops = 2.0 * m * n * k / us / 1e6
    peak = float(peaks.get("bf16_tflops", 2250.0))
    return {
        "workload": "llama3-8b-w4a16-prefill M=4096 K=4096 N=14336 (uint4+zp_int32, g128, fp16)",
        "kernel": "cb200::tc_kernel (" + _lib.VARIANT_NAMES[_lib.load().cb200_last_variant()] + ")",
        "us": us,
        "tflops": tflops,
        "roofline": {
            "bound": "tensor",
            "achieved": tflops,
            "peak": peak,
            "unit": "TFLOP/s",
            "frac": tflops / peak,
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if "bf16_tflops" in peaks else "nominal dense fp16",
            "frac_of_sustained": tflops / float(peaks["bf16_tflops_sustained"]) if "bf16_tflops_sustained" in peaks else None,
        },
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: NCCL all-gather after every GEMM instead of the fused P2P push")
    ap.add_argument("--no-graph", action="store_true", help="time eager Python launches instead of a CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the 70B-FFN (BASELINE config 5) strong-scaling measurement")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch between the GEMVs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
