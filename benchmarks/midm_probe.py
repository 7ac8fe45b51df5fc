"""Probe m_tile / split-K choices of the tcgen05 kernel in the mid-M regime (17 <= M <= 1024)."""
import torch
from common import copies_for_cold_l2, time_rotating
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm

lib = _lib.load()
st = quant_utils.scalar_types.uint4
for name, k, n in (("o", 4096, 4096), ("gate_up", 4096, 28672)):
    wbytes = k * n // 2
    copies = min(16, copies_for_cold_l2(wbytes))
    sets = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, "cuda", seed=i) for i in range(copies)]
    for m in (16, 32, 64, 128, 256, 512):
        x = torch.randn((m, k), device="cuda", dtype=torch.float16)
        out = torch.empty((m, n), device="cuda", dtype=torch.float16)
        best = None
        for variant in (_lib.TC_TS, _lib.TC_SS):
            for mt in (16, 32, 64, 128, 256):
                if mt < min(m, 256) and mt * 8 < m:
                    continue
                if mt > 2 * max(m, 16):
                    continue
                for ks in (1, 2, 4, 8, 16):
                    lib.cb200_set_tc_tuning(mt, ks)
                    try:
                        fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=variant)) for s in sets]
                        us = time_rotating(fns, iters=64, warmup=copies)
                    except Exception as e:  # noqa: BLE001
                        continue
                    finally:
                        lib.cb200_set_tc_tuning(0, 0)
                    row = (us, _lib.VARIANT_NAMES[variant], mt, ks)
                    if best is None or us < best[0]:
                        best = row
                    print(f"  {name} M={m:4d} {row[1]} m_tile={mt:3d} split={ks:2d}: {us:8.2f} us", flush=True)
        # decode kernel for comparison (loops over 16-row passes)
        fns = [(lambda s=s: mixed_precision_gemm(x, s[0], s[1], s[2], 4, 0, 128, out=out, variant=_lib.DECODE_MMA)) for s in sets]
        us_d = time_rotating(fns, iters=64, warmup=copies)
        print(f"BEST {name} M={m:4d}: {best[0]:8.2f} us with {best[1]} m_tile={best[2]} split={best[3]} | decode kernel {us_d:8.2f} us | "
              f"hbm-roof {(wbytes + m * k * 2) / 6545e3:6.2f} us  tensor-roof {2.0 * m * k * n / 1655e6:6.2f} us", flush=True)
    del sets
    torch.cuda.empty_cache()
