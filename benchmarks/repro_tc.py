import pathlib, sys as _s; _s.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import sys, torch
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import mixed_precision_gemm, dequantize_weights
m, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
wname, dname, variant = sys.argv[4], sys.argv[5], int(sys.argv[6])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[dname]
st = getattr(quant_utils.scalar_types, wname)
zp = wname in ("uint4", "uint8")
wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, 128, dtype, "cuda", zero_points=zp, seed=0)
x = torch.randn((m, k), device="cuda", dtype=dtype)
w = dequantize_weights(wq, ws, wzp, st.size_bits, st.bias, 128)
want = x.float() @ w.float()
for i in range(iters):
    out = mixed_precision_gemm(x, wq, ws, wzp, st.size_bits, st.bias, 128, variant=variant)
    torch.cuda.synchronize()
    print(i, "max err", (out.float() - want).abs().max().item(), flush=True)
