import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
import torch
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm
mode = sys.argv[1]
lib = _lib.load()
st = quant_utils.scalar_types.uint4
K, N = (2048, 512) if "big" not in mode else (4096, 28672)
wq, ws, wzp = quant_utils.random_packed_weights(K, N, st, 128, torch.float16, "cuda", seed=1)
x = torch.randn((2, K), device="cuda", dtype=torch.float16)
want = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_STREAM)
flags = torch.zeros(16, dtype=torch.int32, device="cuda")
out = torch.empty_like(want)
main, writer = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
xr = (flags.data_ptr(), 2, flags.data_ptr() + 32)
try:
    if mode.startswith("pre_main"):
        flags[:2] = 1
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
    elif mode.startswith("late_default"):
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
        with torch.cuda.stream(writer):
            torch.cuda._sleep(200_000)
            flags[:2] += 1
    elif mode.startswith("late_main"):
        with torch.cuda.stream(main):
            mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
        with torch.cuda.stream(writer):
            torch.cuda._sleep(200_000)
            flags[:2] += 1
    elif mode.startswith("late_nosleep"):
        with torch.cuda.stream(main):
            mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
        with torch.cuda.stream(writer):
            flags[:2] += 1
    elif mode.startswith("conc_plain"):
        # no chain at all: plain stream GEMVs on main while another kernel runs on writer
        with torch.cuda.stream(writer):
            torch.cuda._sleep(2_000_000)
        with torch.cuda.stream(main):
            for _ in range(20):
                mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, variant=_lib.DECODE_STREAM)
    elif mode.startswith("late_memcpy"):
        host = torch.ones(2, dtype=torch.int32).pin_memory()
        with torch.cuda.stream(main):
            mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
        import time; time.sleep(0.01)
        with torch.cuda.stream(writer):
            flags[:2].copy_(host, non_blocking=True)
    elif mode.startswith("late_hostsleep"):
        with torch.cuda.stream(main):
            mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=xr)
        import time; time.sleep(0.5)
        print("after host sleep, before any other launch", flush=True)
        with torch.cuda.stream(writer):
            flags[:2] += 1
    torch.cuda.synchronize()
    print(mode, "ok", flags.tolist()[:9], torch.equal(out, want), flush=True)
except Exception as e:
    print(mode, "FAILED", repr(e)[:160], flush=True)
