"""GPU parity tests (run with -m gpu on a B200).  Every call goes through the C ABI (ctypes -> libconch_b200.so).

  * dequantised weights: BIT-EXACT against the reference's w_ref (golden fixtures) and against the CPU oracle;
  * GEMM output: rtol = atol = 1e-2 against the oracle's fp32-accumulated result (north_star tolerance), and the
    reference's own looser tolerance against its `torch.matmul(a, w_ref)` fixture.
"""

from __future__ import annotations

import math

import pytest
import torch

import mpgemm_oracle as oracle
from conch_b200 import _lib
from conch_b200.ops.quantization.gemm import dequantize_weights, mixed_precision_gemm
from helpers import ATOL, DTYPES, RTOL, assert_gemm_close, bits, golden_id, golden_index, load_golden, to_cuda

pytestmark = pytest.mark.gpu
INDEX = golden_index()

ALL_VARIANTS = [_lib.DECODE_MMA, _lib.TC_SS, _lib.TC_TS]


def _variants_for(entry: dict) -> list[int]:
    """Every kernel family accepts every fixture (bf16 x int8 runs on the streaming kernel as two nibble planes since round 2)."""
    return ALL_VARIANTS + [_lib.DECODE_STREAM]


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded(built_lib):
    """The tests must exercise the in-tree CUDA library, never a fallback."""
    lib = _lib.load()
    assert lib.cb200_version() == 100
    assert torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    yield
    with open("/proc/self/maps") as f:
        assert "libconch_b200.so" in f.read()


# ------------------------------------------------------------------------------------------------------------
# golden fixtures (produced by the real reference)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("entry", INDEX, ids=golden_id)
def test_dequant_bit_exact_vs_reference_fixture(entry):
    g = load_golden(entry)
    wq, ws, wzp = to_cuda(g["w_q_packed"], g["w_s"], g["w_zp"])
    w = dequantize_weights(wq, ws, wzp, entry["bits"], entry["bias"], entry["group_size"])
    assert w.dtype == DTYPES[entry["dtype"]]
    assert torch.equal(bits(w), bits(g["w_ref"]))
    if wzp is not None:  # vLLM-style uint8 zero-points and activation-dtype zero-points: same bits
        for z in (wzp.to(torch.uint8), wzp.to(w.dtype)):
            w2 = dequantize_weights(wq, ws, z, entry["bits"], entry["bias"], entry["group_size"])
            assert torch.equal(bits(w2), bits(g["w_ref"]))


@pytest.mark.parametrize("entry", INDEX, ids=golden_id)
def test_gemm_vs_reference_fixture(entry):
    g = load_golden(entry)
    x, wq, ws, wzp = to_cuda(g["x"], g["w_q_packed"], g["w_s"], g["w_zp"])
    want = oracle.mixed_precision_gemm_ref(g["x"], g["w_q_packed"], g["w_s"], g["w_zp"], entry["bits"], entry["bias"], entry["group_size"])
    for variant in _variants_for(entry):
        out = mixed_precision_gemm(x, wq, ws, wzp, entry["bits"], entry["bias"], entry["group_size"], variant=variant)
        assert out.shape == (entry["m"], entry["n"]) and out.dtype == x.dtype
        assert_gemm_close(out, want, f"variant {variant}")
        # the reference's own tolerance against its own matmul(a, w_ref) output
        atol = min(5e-2 * math.sqrt(entry["k"]), 1)
        torch.testing.assert_close(out.cpu().float(), g["out_ref_matmul"].float(), rtol=1e-1, atol=atol)


# ------------------------------------------------------------------------------------------------------------
# seeded problems vs the CPU oracle
# ------------------------------------------------------------------------------------------------------------
def _run_case(m, k, n, wname, g, dtype, zp, recipe="llama", seed=0, variant=_lib.AUTO, zp_as=None, out_dtype=None):
    p = oracle.make_problem(m, k, n, wname, g, dtype, zp, seed=seed, recipe=recipe)
    w_zp = p.w_zp
    if w_zp is not None and zp_as is not None:
        w_zp = w_zp.to(zp_as)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, w_zp)
    w = dequantize_weights(wq, ws, wzp, p.wtype.bits, p.wtype.bias, g)
    assert torch.equal(bits(w), bits(p.w_ref)), "dequantised weights must be bit-exact"
    out = mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, g, variant=variant, output_dtype=out_dtype)
    want = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, g, output_dtype=out_dtype)
    assert out.dtype == (out_dtype or dtype)
    assert_gemm_close(out, want, f"{m}x{k}x{n} {wname} g={g} {dtype} zp={zp} variant={variant}")
    return out


@pytest.mark.parametrize("m", [1, 2, 3, 4, 7, 8, 9, 15, 16])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("wname", ["uint4", "uint4b8", "uint8", "uint8b128"])
def test_decode_rows(m, dtype, wname):
    _run_case(m, 1024, 512, wname, 128, dtype, True)


@pytest.mark.parametrize("k,n", [(4096, 4096), (4096, 6144), (14336, 4096), (4096, 28672)], ids=["o", "qkv", "down", "gate_up"])
@pytest.mark.parametrize("m", [1, 4, 16])
def test_llama3_8b_decode_shapes(k, n, m):
    """BASELINE config 2: W4A16, int4 asymmetric (uint4 + zero-points), g=128, fp16."""
    _run_case(m, k, n, "uint4", 128, torch.float16, True)


@pytest.mark.parametrize("g", [64, 128, 256, -1])
@pytest.mark.parametrize("zp", [True, False], ids=["zp", "sym"])
@pytest.mark.parametrize("wname", ["uint4b8", "uint8b128"])
def test_group_sizes(g, zp, wname):
    _run_case(5, 1024, 256, wname, g, torch.bfloat16, zp)


@pytest.mark.parametrize("zp_as", [torch.uint8, torch.float16], ids=["zp_u8", "zp_f16"])
@pytest.mark.parametrize("wname", ["uint4", "uint8"])
def test_zero_point_dtypes(zp_as, wname):
    _run_case(3, 512, 256, wname, 128, torch.float16, True, zp_as=zp_as)


def test_scalar_zero_point():
    p = oracle.make_problem(4, 512, 256, "uint4", 128, torch.float16, False, seed=5)
    zp = torch.tensor([[7]], dtype=torch.int32)
    x, wq, ws, z = to_cuda(p.x, p.w_q_packed, p.w_s, zp)
    out = mixed_precision_gemm(x, wq, ws, z, 4, 0, 128)
    want = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, zp, 4, 0, 128)
    assert_gemm_close(out, want)
    w = dequantize_weights(wq, ws, z, 4, 0, 128)
    assert torch.equal(bits(w), bits(oracle.dequantize(p.w_q_packed, p.w_s, zp, 4, 0, 128, torch.float16)))


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_output_dtype_override(out_dtype):
    _run_case(2, 512, 256, "uint4", 128, torch.float16, True, out_dtype=out_dtype)


def test_non_contiguous_inputs_and_out_buffer():
    p = oracle.make_problem(6, 512, 256, "uint4b8", 128, torch.float16, True, seed=2)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    x_nc = torch.empty((6, 1024), device="cuda", dtype=torch.float16)[:, ::2]
    x_nc.copy_(x)
    assert not x_nc.is_contiguous()
    buf = torch.empty((6, 256), device="cuda", dtype=torch.float16)
    out = mixed_precision_gemm(x_nc, wq, ws, wzp, 4, 8, 128, out=buf)
    assert out.data_ptr() == buf.data_ptr()
    assert_gemm_close(out, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, 4, 8, 128))


def test_m_zero_and_unsupported_shapes():
    p = oracle.make_problem(1, 512, 256, "uint4", 128, torch.float16, True)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    assert mixed_precision_gemm(x[:0], wq, ws, wzp, 4, 0, 128).shape == (0, 256)
    with pytest.raises(ValueError):  # N not a multiple of 64 (outside the reference kernel's domain too)
        mixed_precision_gemm(x, wq[:, :200].contiguous(), ws[:, :200].contiguous(), wzp[:, :200].contiguous(), 4, 0, 128)


def test_split_k_workspace_is_reusable_and_deterministic():
    """Cross-CTA split-K uses self-resetting arrival counters: repeated calls must agree bit for bit."""
    lib = _lib.load()
    p = oracle.make_problem(16, 14336, 256, "uint4", 128, torch.float16, True, seed=3)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    want = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, 4, 0, 128)
    try:
        lib.cb200_set_decode_tuning(0, 0, 8, 0)
        first = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_MMA)
        for _ in range(5):
            again = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_MMA)
            assert torch.equal(again, first)
    finally:
        lib.cb200_set_decode_tuning(0, 0, 0, 0)
    assert_gemm_close(first, want)


@pytest.mark.parametrize("scale_mode", [1, 2], ids=["scale_in_loop", "scale_after"])
@pytest.mark.parametrize("units_per_warp", [0, 1, 2, 3, 5, 64])
def test_decode_tilings(scale_mode, units_per_warp):
    """Stream-K ranges that cut column blocks at every possible place, both scaling modes."""
    lib = _lib.load()
    try:
        lib.cb200_set_decode_tuning(scale_mode, 0, units_per_warp, 0)
        for m in (1, 12):
            _run_case(m, 2048, 512, "uint4", 128, torch.float16, True, variant=_lib.DECODE_MMA)
            _run_case(m, 1024, 256, "uint8b128", 128, torch.bfloat16, True, variant=_lib.DECODE_MMA)
            _run_case(m, 1024, 256, "uint4b8", 64, torch.bfloat16, False, variant=_lib.DECODE_MMA)
    finally:
        lib.cb200_set_decode_tuning(0, 0, 0, 0)


@pytest.mark.parametrize("scale", [2.0**-12, 1e-4, 1e-6])
@pytest.mark.parametrize("variant", [_lib.AUTO, _lib.DECODE_MMA], ids=["auto", "decode_mma"])
def test_default_decode_paths_keep_small_activations(scale, variant):
    """round 1's fused fp16 x int4 form scaled half of the activations by 1/16 in fp16 and lost bits below 2^-10.
    Neither decode kernel does that any more: in both the x16 nibbles accumulate separately and the 1/16 lands on that fp32
    accumulator when the scale group closes."""
    p = oracle.make_problem(4, 4096, 512, "uint4", 128, torch.float16, True, seed=21)
    xs = (p.x.float() * scale).to(torch.float16)
    x, wq, ws, wzp = to_cuda(xs, p.w_q_packed, p.w_s, p.w_zp)
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=variant, output_dtype=torch.float32)
    want = oracle.mixed_precision_gemm_ref(xs, p.w_q_packed, p.w_s, p.w_zp, 4, 0, 128, output_dtype=torch.float32)
    torch.testing.assert_close(out.cpu().float() / scale, want.float() / scale, rtol=1e-2, atol=1e-2)


def test_decode_scale_in_loop_is_the_reference_arithmetic():
    """With the scale applied per weight (mode 1) a one-hot x reproduces the dequantised row bit for bit."""
    lib = _lib.load()
    p = oracle.make_problem(1, 1024, 256, "uint4", 128, torch.float16, True, seed=9)
    _, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    x = torch.zeros((1, 1024), device="cuda", dtype=torch.float16)
    x[0, 517] = 1.0
    try:
        lib.cb200_set_decode_tuning(1, 0, 0, 0)
        out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_MMA)
    finally:
        lib.cb200_set_decode_tuning(0, 0, 0, 0)
    assert torch.equal(bits(out[0]), bits(p.w_ref[517]))


# ------------------------------------------------------------------------------------------------------------
# tcgen05 path
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
@pytest.mark.parametrize("m", [1, 16, 17, 33, 64, 100, 128, 129, 256, 300, 512])
def test_tc_m_sweep(variant, m):
    _run_case(m, 512, 256, "uint4", 128, torch.float16, True, variant=variant)


@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("wname", ["uint4", "uint4b8", "uint8", "uint8b128"])
@pytest.mark.parametrize("zp", [True, False], ids=["zp", "sym"])
def test_tc_formats(variant, dtype, wname, zp):
    _run_case(192, 1024, 384, wname, 128, dtype, zp, variant=variant)


@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
@pytest.mark.parametrize("g", [64, 256, -1])
def test_tc_group_sizes_and_split_k(variant, g):
    lib = _lib.load()
    _run_case(48, 2048, 256, "uint4b8", g, torch.bfloat16, True, variant=variant)
    try:
        lib.cb200_set_tc_tuning(0, 3)
        _run_case(48, 2048, 256, "uint4", g, torch.float16, True, variant=variant)
        lib.cb200_set_tc_tuning(64, 0)
        _run_case(200, 1024, 320, "uint8", g, torch.float16, True, variant=variant, zp_as=torch.uint8)
    finally:
        lib.cb200_set_tc_tuning(0, 0)


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
def test_tc_cluster_multicast(cluster, variant):
    """The activation tile is shared across a CTA cluster by TMA multicast: every cluster size gives the same bits."""
    lib = _lib.load()
    outs = []
    try:
        for cl in (1, cluster):
            lib.cb200_set_tc_cluster(cl)
            outs.append(_run_case(300, 2048, 1024, "uint4", 128, torch.float16, True, variant=variant))
            _run_case(40, 1024, 512, "uint8b128", 128, torch.bfloat16, False, variant=variant)
    finally:
        lib.cb200_set_tc_cluster(0)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
def test_tc_one_hot_rows_reproduce_dequantised_weights(variant):
    """x = one-hot rows => out == W rows bit for bit: the tcgen05 path feeds the SAME dequant bits to the MMA."""
    p = oracle.make_problem(1, 1024, 256, "uint4", 128, torch.float16, True, seed=11)
    _, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    rows = [0, 1, 63, 64, 65, 127, 128, 511, 512, 1023] + list(range(300, 330))
    x = torch.zeros((len(rows), 1024), device="cuda", dtype=torch.float16)
    for i, r in enumerate(rows):
        x[i, r] = 1.0
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=variant)
    assert torch.equal(bits(out), bits(p.w_ref[rows]))


@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
@pytest.mark.parametrize("wname,dtype", [("uint4", torch.float16), ("uint8b128", torch.bfloat16)], ids=["w4_fp16", "w8_bf16"])
def test_tc_repeatability_under_back_to_back_launches(variant, wname, dtype):
    """Race detector: 24 back-to-back launches (long K loop, split-K, many CTAs) must agree bit for bit with the first
    one and with a torch fp32 reference -- a pipeline that reads a stage before TMA has filled it shows up here."""
    from conch_b200 import quant_utils

    st = getattr(quant_utils.scalar_types, wname)
    k, n, m = 4096, 4096, 256
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, st, 128, dtype, "cuda", zero_points=(wname == "uint4"), seed=5)
    x = torch.randn((m, k), device="cuda", dtype=dtype)
    outs = [mixed_precision_gemm(x, wq, ws, wzp, st.size_bits, st.bias, 128, variant=variant) for _ in range(24)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    w = dequantize_weights(wq, ws, wzp, st.size_bits, st.bias, 128)
    ref = (x.float() @ w.float()).to(dtype)
    torch.testing.assert_close(outs[0].float(), ref.float(), rtol=RTOL, atol=ATOL if dtype == torch.float16 else 0.5)


@pytest.mark.parametrize("variant", [_lib.TC_SS, _lib.TC_TS], ids=["ss", "ts"])
def test_tc_prefill_shape(variant):
    """A slice of BASELINE config 3 (M=4096 would take the CPU oracle too long): M=512, K=4096, N=1024."""
    _run_case(512, 4096, 1024, "uint4", 128, torch.float16, True, variant=variant)


# ------------------------------------------------------------------------------------------------------------
# the reference's own test matrix (tests/mixed_precision_gemm_test.py:36-77), re-run against this op
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize(("m_dim", "k_dim", "n_dim"), [(128, 256, 128), (1024, 1024, 1024)])
@pytest.mark.parametrize("weight_dtype", ["uint4b8", "uint8b128", "uint4", "uint8"])
@pytest.mark.parametrize("use_zero_points", [True, False])
@pytest.mark.parametrize("input_dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_reference_matrix(m_dim, k_dim, n_dim, weight_dtype, use_zero_points, input_dtype):
    p = oracle.make_problem(m_dim, k_dim, n_dim, weight_dtype, 128, input_dtype, use_zero_points, seed=0, recipe="reference")
    a, w_q_packed, w_s, w_zp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    output_ref = oracle.reference_test_expression(p.x, p.w_ref)  # torch.matmul(a, w_ref), as the reference test does
    output = mixed_precision_gemm(a, w_q_packed, w_s, w_zp, p.wtype.bits, p.wtype.bias, 128)
    atol = min(5e-2 * math.sqrt(k_dim), 1)
    torch.testing.assert_close(output.cpu(), output_ref, rtol=1e-1, atol=atol)  # the reference's tolerance
    want = oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, 128)
    assert_gemm_close(output, want)  # the north_star tolerance


# ------------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE's full sizes (no CPU oracle needed)
# ------------------------------------------------------------------------------------------------------------
def _torch_dequant_gpu(wq, ws, wzp, bits_, bias, g):
    """Plain-PyTorch restatement on the GPU (integer unpack is exact; one rounding in the multiply)."""
    pf = 32 // bits_
    kp, n = wq.shape
    shifts = (torch.arange(pf, device=wq.device, dtype=torch.int32) * bits_).view(1, pf, 1)
    q = ((wq.view(kp, 1, n) >> shifts) & ((1 << bits_) - 1)).reshape(kp * pf, n)
    v = q - bias
    if wzp is not None:
        v = v - wzp.to(torch.int32).repeat_interleave(g, dim=0)
    return v.to(ws.dtype) * ws.repeat_interleave(g, dim=0)


@pytest.mark.parametrize("k,n", [(4096, 28672), (14336, 4096), (8192, 28672)], ids=["8b_gate_up", "8b_down", "70b_ffn"])
def test_full_size_row_selection_and_dequant(k, n):
    """x = one-hot rows  =>  out[i, :] == W[k_i, :] exactly; and the dequant kernel equals a torch restatement."""
    from conch_b200 import quant_utils

    wq, ws, wzp = quant_utils.random_packed_weights(k, n, quant_utils.scalar_types.uint4, 128, torch.float16, "cuda", seed=1)
    w = dequantize_weights(wq, ws, wzp, 4, 0, 128)
    assert torch.equal(w, _torch_dequant_gpu(wq, ws, wzp, 4, 0, 128))
    rows = [0, 1, 7, 8, 127, 128, k // 2 + 3, k - 1]
    for m in (1, 8):
        x = torch.zeros((m, k), device="cuda", dtype=torch.float16)
        for i in range(m):
            x[i, rows[i]] = 1.0
        out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
        assert torch.equal(out, w[rows[:m]])
    # linearity in x (fp32 accumulate): f(2x) == 2 f(x) exactly, f(x1) + f(x2) ~= f(x1 + x2)
    gen = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn((4, k), generator=gen, device="cuda", dtype=torch.float16)
    f1 = mixed_precision_gemm(x1, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
    f2 = mixed_precision_gemm(x1 * 2, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
    # the default (fused) decode path pre-scales some activations by 1/16: exact except for fp16 subnormals
    torch.testing.assert_close(f2, f1 * 2, rtol=1e-3, atol=1e-3)
    lib = _lib.load()
    try:  # per-weight dequant mode: scaling x by 2 is exact end to end
        lib.cb200_set_decode_tuning(1, 0, 0, 0)
        g1 = mixed_precision_gemm(x1, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
        g2 = mixed_precision_gemm(x1 * 2, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
    finally:
        lib.cb200_set_decode_tuning(0, 0, 0, 0)
    assert torch.equal(g2, g1 * 2)
    # fused keeps W unrounded, per-weight mode rounds each W to fp16 like the reference: both inside the 1e-2 budget
    torch.testing.assert_close(f1, g1, rtol=5e-3, atol=5e-3)
    ref = x1.float() @ w.float()
    torch.testing.assert_close(f1, ref, rtol=RTOL, atol=ATOL)


def test_batch_rows_equal_single_rows():
    """Row m of an M=24 call against the M=1 call on that row.  AUTO sends M=24 to the tcgen05 kernel, which rounds W to fp16
    exactly like the reference, and M=1 to the streaming kernel, which keeps s*(q-zp) unrounded in fp32: the two differ by
    sum_k x_k * delta_k with delta_k the fp16 rounding error of W[k,n], |delta_k| <= 2^-11 |W| (half an ulp), independent and
    close to uniform, so rms(delta) <= 2^-10 max|W| / sqrt(12).  The tolerance is 6 sigma of that sum -- derived, not picked
    (it evaluates to ~4e-3 here; the observed maximum is ~1.2e-3 over 18k outputs)."""
    from conch_b200 import quant_utils

    k, n = 4096, 6144
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, quant_utils.scalar_types.uint4, 128, torch.float16, "cuda", seed=2)
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((24, k), generator=gen, device="cuda", dtype=torch.float16)
    w = dequantize_weights(wq, ws, wzp, 4, 0, 128)
    sigma = math.sqrt(k) * float(x.float().pow(2).mean().sqrt()) * (2.0**-10 * float(w.abs().max())) / math.sqrt(12.0)
    full = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
    assert _lib.load().cb200_last_variant() == _lib.TC_TS
    for i in (0, 5, 15):
        one = mixed_precision_gemm(x[i : i + 1].contiguous(), wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32)
        assert _lib.load().cb200_last_variant() == _lib.DECODE_STREAM
        torch.testing.assert_close(one[0], full[i], rtol=0.0, atol=6.0 * sigma)


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU epilogue (cb200_mpgemm_push) exercised on ONE device: the "peers" are just other local buffers
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 3, 8])
@pytest.mark.parametrize("k,n", [(1024, 512), (4096, 6144)])
def test_push_epilogue_writes_every_peer_buffer(m, k, n):
    from conch_b200 import quant_utils

    wq, ws, wzp = quant_utils.random_packed_weights(k, n, quant_utils.scalar_types.uint4, 128, torch.float16, "cuda", seed=11)
    x = torch.randn((m, k), device="cuda", dtype=torch.float16)
    plain = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    peers = [torch.full((m, n), float("nan"), device="cuda", dtype=torch.float16) for _ in range(3)]
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, peer_out_ptrs=[t.data_ptr() for t in peers])
    torch.cuda.synchronize()
    assert torch.equal(out, plain)
    for t in peers:
        assert torch.equal(t, plain)


def test_push_signal_counts_and_wait_is_replay_safe():
    """cb200_mpgemm_push_signal: after the whole result has been pushed the kernel bumps the local and the peer counters;
    cb200_wait_flags keeps its own count on the device, so a (GEMV, wait) pair replays from a CUDA graph.  Single GPU: the
    "peers" are buffers of the same device."""
    lib = _lib.load()
    p = oracle.make_problem(3, 2048, 1024, "uint4", 128, torch.float16, True, seed=9)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    want = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    peers = [torch.zeros_like(want) for _ in range(2)]
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")  # [0..2]: counters of "rank" 0 (local), 1, 2; [8]: the wait's own count
    out = torch.empty_like(want)
    sig = ([flags.data_ptr() + 4, flags.data_ptr() + 8], flags.data_ptr())

    def call():
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, peer_out_ptrs=[t.data_ptr() for t in peers], signal=sig)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.raise_for_status(lib.cb200_wait_flags(flags.data_ptr(), 3, flags.data_ptr() + 32, stream))

    call()
    torch.cuda.synchronize()
    assert flags[:3].tolist() == [1, 1, 1] and int(flags[8]) == 1
    for t in peers + [out]:
        assert torch.equal(bits(t), bits(want))
    call()
    torch.cuda.synchronize()
    assert flags[:3].tolist() == [2, 2, 2] and int(flags[8]) == 2
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        for _ in range(5):
            g.replay()
    torch.cuda.synchronize()
    assert flags[:3].tolist() == [7, 7, 7] and int(flags[8]) == 7
    with pytest.raises(NotImplementedError):  # the flags are raised by the decode kernels only
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out, variant=_lib.TC_TS, peer_out_ptrs=[peers[0].data_ptr()], signal=sig)


def test_push_epilogue_is_decode_only():
    from conch_b200 import _lib, quant_utils

    wq, ws, wzp = quant_utils.random_packed_weights(512, 256, quant_utils.scalar_types.uint4, 128, torch.float16, "cuda", seed=12)
    x = torch.randn((64, 512), device="cuda", dtype=torch.float16)
    peer = torch.empty((64, 256), device="cuda", dtype=torch.float16)
    with pytest.raises(NotImplementedError):
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, peer_out_ptrs=[peer.data_ptr()])  # AUTO -> tcgen05 kernel at M=64
    with pytest.raises(ValueError):
        mixed_precision_gemm(x[:1].contiguous(), wq, ws, wzp, 4, 0, 128, peer_out_ptrs=[peer.data_ptr()] * 8)  # > 7 peers
    # forcing the decode kernel works for any M (16 rows per pass)
    got = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.DECODE_MMA, peer_out_ptrs=[peer.data_ptr()])
    torch.cuda.synchronize()
    assert torch.equal(peer, got)


# ------------------------------------------------------------------------------------------------------------------
# persistent tcgen05 variant (CB200_TC2_TS): several tiles per CTA, double-buffered TMEM accumulators
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n", [(1, 512, 256), (128, 512, 256), (300, 1024, 320), (1000, 2048, 1024), (4096, 1024, 6144)])
@pytest.mark.parametrize("dtype,wname,zp,g", [(torch.float16, "uint4", True, 128), (torch.bfloat16, "uint4b8", False, 64),
                                              (torch.float16, "uint8", True, -1), (torch.bfloat16, "uint8b128", False, 256)])
def test_tc2_persistent_variant(m, k, n, dtype, wname, zp, g):
    # (4096, 1024, 6144) is 32 x 48 = 1536 tiles: ~10 tiles per CTA through both accumulator buffers
    _run_case(m, k, n, wname, g, dtype, zp, variant=_lib.TC2_TS)


def test_tc2_matches_tc_ts_bit_for_bit_and_repeats():
    from conch_b200 import quant_utils

    k, n = 4096, 4096
    wq, ws, wzp = quant_utils.random_packed_weights(k, n, quant_utils.scalar_types.uint4, 128, torch.float16, "cuda", seed=21)
    x = torch.randn((2048, k), device="cuda", dtype=torch.float16)
    a = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.TC2_TS)
    b = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.TC2_TS)
    ref = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=_lib.TC_TS)
    assert torch.equal(a, b)
    # same K order, one accumulator per tile in both kernels (TC_TS splits K over two accumulators only for m_tile <= 128)
    assert torch.equal(a, ref)


@pytest.mark.parametrize("m,k,n,dtype,wname", [(512, 1024, 10240, torch.float16, "uint4"), (300, 2048, 10176, torch.bfloat16, "uint4b8"),
                                               (2048, 512, 2560, torch.float16, "uint8")])
def test_tc_split_of_the_partial_last_wave(m, k, n, dtype, wname):
    """More tiles than SMs with a small remainder: the last wave's tiles are split along K (work items after the whole tiles)."""
    _run_case(m, k, n, wname, 128, dtype, wname != "uint4b8", variant=_lib.TC_TS)
    _run_case(m, k, n, wname, 128, dtype, wname != "uint4b8", variant=_lib.TC_SS)


# ------------------------------------------------------------------------------------------------------------------
# smallest / odd shapes of the domain: one K block, column-block counts that are not multiples of 4, ragged M
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n", [(1, 64, 64), (3, 128, 192), (8, 192, 64), (5, 320, 448), (16, 64, 64), (17, 64, 64), (33, 128, 320),
                                   (130, 192, 192), (257, 64, 128), (300, 448, 64)])
@pytest.mark.parametrize("dtype,wname,zp,g", [(torch.float16, "uint4", True, 64), (torch.bfloat16, "uint4b8", False, -1),
                                              (torch.float16, "uint8b128", False, 64), (torch.bfloat16, "uint8", True, -1)])
def test_smallest_and_odd_shapes(m, k, n, dtype, wname, zp, g):
    _run_case(m, k, n, wname, g, dtype, zp)
    if m <= 16:
        _run_case(m, k, n, wname, g, dtype, zp, variant=_lib.DECODE_MMA)
    _run_case(m, k, n, wname, g, dtype, zp, variant=_lib.TC_TS)
    _run_case(m, k, n, wname, g, dtype, zp, variant=_lib.TC2_TS)


@pytest.mark.parametrize("m", [1, 2])
@pytest.mark.parametrize("dtype,wname,zp", [(torch.float16, "uint4", True), (torch.float16, "uint8", True), (torch.bfloat16, "uint4b8", False),
                                            (torch.float16, "uint8b128", False)])
@pytest.mark.parametrize("g", [64, 128, 256, -1])
def test_decode_group_sums_prologue(m, dtype, wname, zp, g):
    """M <= 2 on long slabs: the CTA computes the group sums of x once (and rewrites the slab in operand order) instead of
    an all-ones MMA per operand pair.  One split forces 16..32 batches per warp, i.e. that path."""
    lib = _lib.load()
    try:
        lib.cb200_set_decode_tuning(2, 0, 1, 0)
        _run_case(m, 2048, 320, wname, g, dtype, zp, variant=_lib.DECODE_MMA)
        lib.cb200_set_decode_tuning(2, 0, 2, 0)
        _run_case(m, 4096, 192, wname, g, dtype, zp, variant=_lib.DECODE_MMA, zp_as=torch.uint8 if zp else None)
    finally:
        lib.cb200_set_decode_tuning(0, 0, 0, 0)


def test_repeat_calls_take_the_fast_path_and_stay_correct():
    """The op remembers what it validated about a set of weight tensors (conch_b200/ops/quantization/gemm.py, _fast_entries):
    repeat calls skip the metadata pass.  Same results, same errors, and nothing is reused for different tensors."""
    import gc

    from conch_b200.ops.quantization import gemm as og

    p = oracle.make_problem(3, 512, 256, "uint4", 128, torch.float16, True, seed=11)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    y1 = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    assert id(wq) in og._fast_entries
    y2 = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128)
    out = torch.empty_like(y1)
    y3 = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=out)
    assert y3 is out and torch.equal(bits(y1), bits(y2)) and torch.equal(bits(y1), bits(y3))
    assert_gemm_close(y2, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, 4, 0, 128))
    assert torch.equal(bits(mixed_precision_gemm(x[:1].contiguous(), wq, ws, wzp, 4, 0, 128)), bits(y1[:1]))  # another M
    y4 = mixed_precision_gemm(x, wq, ws * 2, wzp, 4, 0, 128)  # other scales with the same packed tensor: not the remembered entry
    torch.testing.assert_close(y4.float(), 2 * y1.float(), rtol=2e-3, atol=2e-3)
    with pytest.raises(ValueError):
        mixed_precision_gemm(x[:, :256].contiguous(), wq, ws, wzp, 4, 0, 128)  # K mismatch: full validation words the error
    with pytest.raises(ValueError):
        mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, out=torch.empty((3, 256), device="cuda", dtype=torch.float32))
    with pytest.raises(ValueError):
        mixed_precision_gemm(x, wq, ws[:2], wzp, 4, 0, 128)
    assert mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, output_dtype=torch.float32).dtype == torch.float32  # keyword: full path
    key = id(wq)
    del wq, y1, y2, y3, y4
    gc.collect()
    assert key not in og._fast_entries
