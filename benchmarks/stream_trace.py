"""Timeline of a CHAIN of streaming decode kernels (cb200_debug_set_stream_trace): where the time between two consecutive
GEMVs of a decode token goes, and how much of kernel i+1 overlaps kernel i under programmatic dependent launch.

    python benchmarks/stream_trace.py [--layers 3] [--m 1] [--no-pdl] [--no-graph]

Stamp slots (per CTA, %globaltimer ns): 0 entry, 1 barriers initialised, 2 producer: first ring requested (static data),
3 producer: griddepcontrol.wait returned, 4 producer: last stage requested, 5 consumer warp 0: griddepcontrol.wait returned,
6 consumer warp 0: first stage landed, 7 consumer warp 0: last column block finished (before its publish), 8 consumer warp 0: done,
9 SM id.
"""

from __future__ import annotations

import argparse

import torch

import common  # noqa: F401
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

PROJ = (("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--variant", type=int, default=_lib.DECODE_TCS)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda")
    st = quant_utils.scalar_types.uint4
    kernels = []
    for layer in range(args.layers + 8):  # + rotation so that nothing is L2-resident
        for name, k, n in PROJ:
            w = quant_utils.random_packed_weights(k, n, st, 128, torch.float16, dev, seed=4 * layer + len(kernels))
            kernels.append((name, k, n, w))
    xs = {k: torch.randn((args.m, k), device=dev, dtype=torch.float16) for k in (4096, 14336)}
    outs = [torch.empty((args.m, n), device=dev, dtype=torch.float16) for _, _, n, _ in kernels]
    traces = [torch.zeros((1024, 32), dtype=torch.int64, device=dev) for _ in kernels]
    lib.cb200_set_pdl(0 if args.no_pdl else 1)

    def run(traced: bool):
        for i, (name, k, n, w) in enumerate(kernels):
            lib.cb200_debug_set_stream_trace(traces[i].data_ptr() if traced else None)
            mixed_precision_gemm(xs[k], w[0], w[1], w[2], 4, 0, 128, out=outs[i], variant=args.variant)
        lib.cb200_debug_set_stream_trace(None)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run(False)
        s.synchronize()
        if args.no_graph:
            run(True)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                run(True)
            for _ in range(3):
                g.replay()
        s.synchronize()
    lib.cb200_set_pdl(0)

    tr = [t.cpu() for t in traces]
    t0 = min(int(t[:, 0][t[:, 0] > 0].min()) for t in tr)

    def stat(t, c):
        v = t[:, c][t[:, c] > 0] - t0
        if len(v) == 0:
            return "      -      "
        v = v.float() / 1e3
        return f"{v.min():6.1f}/{v.median():6.1f}/{v.max():6.1f}"

    print(f"chain of {len(kernels)} GEMVs, M={args.m}, pdl={not args.no_pdl}, graph={not args.no_graph}; us since the first entry, min/median/max over CTAs")
    print(f"{'#':>3s} {'layer':8s} {'entry':>20s} {'ring requested':>20s} {'pdl wait (prod)':>20s} {'first stage':>20s} {'last cb done':>20s} {'done':>20s}   dur   gap")
    prev_end = None
    skip = 8 * len(PROJ) // 2
    for i, t in enumerate(tr):
        end = float((t[:, 8][t[:, 8] > 0] - t0).max()) / 1e3
        start = float((t[:, 0][t[:, 0] > 0] - t0).min()) / 1e3
        gap = "" if prev_end is None else f"{end - prev_end:6.2f}"
        if i >= skip and i < skip + 4 * args.layers:
            print(f"{i:3d} {kernels[i][0]:8s} {stat(t, 0):>20s} {stat(t, 2):>20s} {stat(t, 3):>20s} {stat(t, 6):>20s} {stat(t, 7):>20s} {stat(t, 8):>20s} {end - start:6.2f} {gap}")
        prev_end = end
    # phase accounting of the tcgen05 kernel (SM cycles, set 0's roles, median over CTAs)
    names = ["issuer: waits", "issuer: issue", "", "", "", "", "x warp: waits", "x warp: permute + sums", "x warp: group end + signal", "", "", "",
             "dequant: fold", "dequant: wait full", "dequant: load words + meta", "dequant: wait mma_done", "dequant: unpack + tcgen05.st", ""]
    for i in range(skip, skip + 4):
        t = tr[i]
        used = t[:, 0] > 0
        if int(t[used][:, 10:28].abs().sum()) == 0:
            continue
        nst = kernels[i][1] * kernels[i][2] // 2 // 8192 / max(1, int(used.sum()))
        print(f"  phases of {kernels[i][0]} ({nst:.1f} stages per CTA):")
        for k, nm in enumerate(names):
            if nm:
                v = t[used][:, 10 + k].float()
                print(f"    {nm:32s} median {v.median():9.0f} cycles  ({v.median() / max(nst / 2, 1):7.0f} per own stage)")
    sms = {int(v) for t in tr[skip : skip + 4] for v in t[:, 9].tolist()}
    print(f"SMs used by the consumers of 4 consecutive kernels: {len(sms)}")


if __name__ == "__main__":
    main()
