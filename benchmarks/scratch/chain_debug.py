import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2])); sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2] / "oracle")); sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2] / "tests"))
import torch
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm
lib = _lib.load()
st = quant_utils.scalar_types.uint4
wq, ws, wzp = quant_utils.random_packed_weights(2048, 512, st, 128, torch.float16, "cuda", seed=1)
x = torch.randn((2, 2048), device="cuda", dtype=torch.float16)
want = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_STREAM)
flags = torch.zeros(16, dtype=torch.int32, device="cuda")
out = torch.empty_like(want)
torch.cuda.synchronize(); print("plain ok", flush=True)
# (a) flags raised beforehand
flags[:2] = 1
torch.cuda.synchronize()
try:
    mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, x_ready=(flags.data_ptr(), 2, flags.data_ptr() + 32))
    torch.cuda.synchronize(); print("(a) ok", flags.tolist(), torch.equal(out, want), flush=True)
except Exception as e:
    print("(a) FAILED", repr(e)[:300], flush=True); sys.exit(0)
# (b) flags raised by a host-side write from another stream after a delay
xbuf = torch.zeros_like(x)
main, writer = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(main):
    mixed_precision_gemm(xbuf, wq, ws, wzp, 4, 0, 128, out=out, x_ready=(flags.data_ptr(), 2, flags.data_ptr() + 32))
print("(b) launched", flush=True)
try:
    with torch.cuda.stream(writer):
        torch.cuda._sleep(1_000_000)
        print("(b) sleep launched", flush=True)
        peer = (ctypes.c_void_p * 1)(xbuf.data_ptr()); pflag = (ctypes.c_void_p * 1)(flags.data_ptr() + 4)
        _lib.raise_for_status(lib.cb200_push(x.data_ptr(), x.numel() * 2, peer, 1, pflag, flags.data_ptr(), writer.cuda_stream))
    torch.cuda.synchronize(); print("(b) ok", flags.tolist(), torch.equal(out, want), flush=True)
except Exception as e:
    print("(b) FAILED", repr(e)[:300], flush=True)
