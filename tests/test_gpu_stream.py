"""GPU parity tests of the streaming ring decode kernel (CB200_DECODE_STREAM, AUTO's choice for M <= 8), through the C ABI.

The kernel never forms the rounded W (scale / zero-point are applied per scale group in fp32), so the GEMM bar is the
north_star tolerance (rtol = atol = 1e-2 against the fp32-accumulated oracle); one-hot activations still have to reproduce
the reference's dequantised weights BIT FOR BIT.  The overlap protocol (programmatic dependent launch: the next kernel
prefetches its weights while this one runs) is exercised by back-to-back chains, eagerly and from a CUDA graph.
"""

from __future__ import annotations

import pytest
import torch

import mpgemm_oracle as oracle
from conch_b200 import _lib, quant_utils
from conch_b200.ops.quantization.gemm import dequantize_weights, mixed_precision_gemm
from helpers import assert_gemm_close, bits, to_cuda

pytestmark = pytest.mark.gpu
STREAM = _lib.DECODE_STREAM


@pytest.fixture(scope="module", autouse=True)
def _native_library_loaded(built_lib):
    lib = _lib.load()
    assert torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    yield
    lib.cb200_set_pdl(0)
    lib.cb200_set_stream_tuning(0, 0)
    with open("/proc/self/maps") as f:
        assert "libconch_b200.so" in f.read()


def _case(m, k, n, wname, g, dtype, zp, seed=0, zp_as=None, out_dtype=None, recipe="llama", variant=STREAM, x_scale=None):
    p = oracle.make_problem(m, k, n, wname, g, dtype, zp, seed=seed, recipe=recipe)
    w_zp = p.w_zp if (p.w_zp is None or zp_as is None) else p.w_zp.to(zp_as)
    xh = p.x if x_scale is None else (p.x.float() * x_scale).to(dtype)
    x, wq, ws, wzp = to_cuda(xh, p.w_q_packed, p.w_s, w_zp)
    out = mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, g, variant=variant, output_dtype=out_dtype)
    if variant != _lib.AUTO:
        assert _lib.load().cb200_last_variant() == variant
    want = oracle.mixed_precision_gemm_ref(xh, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, g, output_dtype=out_dtype)
    return out, want, p


@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 9, 13, 16])
@pytest.mark.parametrize("dtype,wname", [(torch.float16, "uint4"), (torch.float16, "uint4b8"), (torch.float16, "uint8"),
                                         (torch.float16, "uint8b128"), (torch.bfloat16, "uint4"), (torch.bfloat16, "uint4b8"),
                                         (torch.bfloat16, "uint8"), (torch.bfloat16, "uint8b128")],
                         ids=["f16-u4", "f16-u4b8", "f16-u8", "f16-u8b128", "bf16-u4", "bf16-u4b8", "bf16-u8", "bf16-u8b128"])  # fmt: skip
def test_rows_and_formats(m, dtype, wname):
    out, want, _ = _case(m, 1024, 512, wname, 128, dtype, True)
    assert_gemm_close(out, want, f"M={m} {wname} {dtype}")


def test_auto_dispatch_and_unsupported_formats():
    """AUTO: the streaming kernel up to 16 rows where it has a stage layout for the problem, round 1's register-window kernel
    for the rest up to 8 rows (group sizes that straddle a stage), tcgen05 above."""
    lib = _lib.load()

    def auto(m, k, n, wname, g, dtype):
        p = oracle.make_problem(m, k, n, wname, g, dtype, True, seed=0)
        x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
        out = mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, g)
        assert_gemm_close(out, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, p.wtype.bits, p.wtype.bias, g))
        return lib.cb200_last_variant()

    assert auto(1, 512, 256, "uint4", 128, torch.float16) == STREAM
    assert auto(8, 512, 256, "uint4", 128, torch.float16) == STREAM
    assert auto(9, 512, 256, "uint4", 128, torch.float16) == STREAM  # 16-row variant
    assert auto(16, 512, 256, "uint4", 128, torch.float16) == STREAM
    assert auto(17, 512, 256, "uint4", 128, torch.float16) == _lib.TC_TS
    assert auto(8, 512, 256, "uint8", 128, torch.bfloat16) == STREAM
    assert auto(12, 512, 256, "uint8", 128, torch.float16) == _lib.TC_TS  # 8-bit weights: tcgen05 above 8 rows
    assert auto(12, 768, 256, "uint4", 192, torch.float16) == _lib.TC_TS  # no streaming stage layout: tcgen05 above 8 rows
    assert auto(4, 512, 256, "uint8", 128, torch.float16) == STREAM
    assert auto(4, 512, 256, "uint4b8", -1, torch.bfloat16) == STREAM
    assert auto(4, 512, 256, "uint8", 128, torch.bfloat16) == STREAM  # (two nibble planes: a byte does not fit a bf16 significand)
    assert auto(4, 768, 256, "uint4", 192, torch.float16) == _lib.DECODE_MMA   # 6 quads per group: straddles the 4-quad stages
    for wname, g, dtype, k in (("uint4", 192, torch.float16, 768),):
        p = oracle.make_problem(4, k, 256, wname, g, dtype, True, seed=0)
        x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
        with pytest.raises(NotImplementedError):
            mixed_precision_gemm(x, wq, ws, wzp, p.wtype.bits, p.wtype.bias, g, variant=STREAM)


@pytest.mark.parametrize("g", [64, 128, 192, 256, 384, -1])
@pytest.mark.parametrize("zp", [True, False], ids=["zp", "sym"])
@pytest.mark.parametrize("dtype,wname", [(torch.float16, "uint4"), (torch.bfloat16, "uint4b8"), (torch.float16, "uint8b128")],
                         ids=["f16-u4", "bf16-u4b8", "f16-u8b128"])  # fmt: skip
def test_group_sizes(g, zp, dtype, wname):
    k = 1536 if g in (192, 384) else 1024  # K % g == 0
    if g == 192 and wname.startswith("uint4"):
        pytest.skip("int4 with 192-k groups (6 quads of packed rows) straddles the kernel's 4-quad stages: CB200_DECODE_MMA's")
    out, want, _ = _case(5, k, 256, wname, g, dtype, zp)
    assert_gemm_close(out, want, f"g={g}")


@pytest.mark.parametrize("n", [64, 192, 320, 1088])  # N % 128 == 64: the last 128-column block is half empty
@pytest.mark.parametrize("m", [1, 11])
def test_ragged_last_column_block(n, m):
    out, want, _ = _case(m, 512, n, "uint4", 128, torch.float16, True)
    assert_gemm_close(out, want, f"N={n}")


@pytest.mark.parametrize("zp_as", [torch.uint8, torch.float16], ids=["zp_u8", "zp_f16"])
@pytest.mark.parametrize("wname", ["uint4", "uint8"])
def test_zero_point_dtypes(zp_as, wname):
    out, want, _ = _case(3, 512, 256, wname, 128, torch.float16, True, zp_as=zp_as)
    assert_gemm_close(out, want)


def test_scalar_zero_point_and_output_dtypes():
    p = oracle.make_problem(4, 512, 256, "uint4", 128, torch.float16, False, seed=5)
    zp = torch.tensor([[7]], dtype=torch.int32)
    x, wq, ws, z = to_cuda(p.x, p.w_q_packed, p.w_s, zp)
    out = mixed_precision_gemm(x, wq, ws, z, 4, 0, 128, variant=STREAM)
    assert_gemm_close(out, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, zp, 4, 0, 128))
    for od in (torch.float32, torch.bfloat16):
        out, want, _ = _case(2, 512, 256, "uint4", 128, torch.float16, True, out_dtype=od)
        assert out.dtype == od
        assert_gemm_close(out, want)


@pytest.mark.parametrize("k,n", [(4096, 4096), (4096, 6144), (14336, 4096), (4096, 28672)], ids=["o", "qkv", "down", "gate_up"])
@pytest.mark.parametrize("m", [1, 4, 16])
def test_llama3_8b_decode_shapes(k, n, m):
    """BASELINE config 2: W4A16, int4 asymmetric (uint4 + zero-points), g=128, fp16, full size."""
    out, want, _ = _case(m, k, n, "uint4", 128, torch.float16, True)
    assert_gemm_close(out, want, f"{m}x{k}x{n}")


@pytest.mark.parametrize("wname,g,dtype", [("uint8", 128, torch.float16), ("uint8b128", 128, torch.float16),
                                           ("uint4", -1, torch.bfloat16), ("uint4b8", -1, torch.bfloat16),
                                           ("uint8", 128, torch.bfloat16), ("uint8b128", -1, torch.bfloat16)])  # fmt: skip
@pytest.mark.parametrize("m", [1, 16])
def test_config4_formats_at_llama_size(wname, g, dtype, m):
    """BASELINE config 4 (W8A16, per-channel int4 with bf16) on the 4096 x 4096 projection."""
    out, want, _ = _case(m, 4096, 4096, wname, g, dtype, wname in ("uint8", "uint4"))
    assert_gemm_close(out, want)


@pytest.mark.parametrize("dtype,wname,zp", [(torch.float16, "uint4", True), (torch.float16, "uint4b8", False),
                                            (torch.float16, "uint8", True), (torch.bfloat16, "uint4", True),
                                            (torch.bfloat16, "uint8", True), (torch.bfloat16, "uint8b128", False)])  # fmt: skip
def test_one_hot_rows_reproduce_the_reference_weights_bit_for_bit(dtype, wname, zp):
    """x = e_k selects row k of W: the per-group fp32 correction s*q - s*zb is exact, so the single rounding to the output
    dtype is the reference's (conch/kernels/quantization/gemm.py:205-216)."""
    k_dim, n = 2048, 384
    p = oracle.make_problem(1, k_dim, n, wname, 128, dtype, zp, seed=3)
    rows = [0, 1, 7, 8, 127, 128, 129, 1000, 1033, 2047, 5, 640, 77, 1291, 255, 256]
    x = torch.zeros((16, k_dim), dtype=dtype)
    for i, r in enumerate(rows):
        x[i, r] = 1.0
    xd, wq, ws, wzp = to_cuda(x, p.w_q_packed, p.w_s, p.w_zp)
    for m in (1, 8, 16):
        out = mixed_precision_gemm(xd[:m].contiguous(), wq, ws, wzp, p.wtype.bits, p.wtype.bias, 128, variant=STREAM)
        assert torch.equal(bits(out), bits(p.w_ref[rows[:m]])), f"M={m}"


@pytest.mark.parametrize("scale", [2.0**-12, 2.0**-20, 1e-4, 1e-6])
def test_small_magnitude_activations_keep_their_bits(scale):
    """Round 1 pre-multiplied half of the fp16 activations by 1/16 (lost bits below 2^-10, flushed below 2^-21).  The streaming
    kernel scales the odd-nibble ACCUMULATOR in fp32 instead: relative accuracy must not depend on the magnitude of x."""
    out, want, _ = _case(4, 4096, 512, "uint4", 128, torch.float16, True, out_dtype=torch.float32, x_scale=scale)
    torch.testing.assert_close(out.cpu().float() / scale, want.float() / scale, rtol=1e-2, atol=1e-2)


def test_reference_recipe_at_llama_k():
    """The reference test's input recipe 10*(U-0.3) (large mean => cancellation between s*G and s*zb*X) at K = 14336."""
    out, want, _ = _case(2, 14336, 256, "uint4", 128, torch.float16, True, recipe="reference", out_dtype=torch.float32)
    torch.testing.assert_close(out.cpu().float(), want.float(), rtol=1e-2, atol=1e-2)


def test_more_than_16_rows_runs_in_passes():
    out, want, _ = _case(37, 1024, 256, "uint4", 128, torch.float16, True)
    assert_gemm_close(out, want)
    assert _lib.load().cb200_last_launch_count() == 3


def test_batch_rows_equal_single_rows_bit_for_bit():
    """Row m of an M-row call sees exactly the arithmetic of a 1-row call (same stages, same MMA k order, same slot order in
    the split-K sum): the results must be IDENTICAL, not merely close."""
    p = oracle.make_problem(8, 4096, 1024, "uint4", 128, torch.float16, True, seed=11)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    full = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=STREAM)
    for m in range(8):
        one = mixed_precision_gemm(x[m : m + 1].contiguous(), wq, ws, wzp, 4, 0, 128, variant=STREAM)
        assert torch.equal(bits(one[0]), bits(full[m])), f"row {m}"


@pytest.mark.parametrize("ring_kb,ctas", [(24, 0), (0, 7), (0, 148), (60, 33), (200, 0)])
def test_ring_depth_and_grid_do_not_change_the_result(ring_kb, ctas):
    lib = _lib.load()
    try:
        lib.cb200_set_stream_tuning(ring_kb, ctas)
        for m, k, n, wname, dtype in ((1, 2048, 1024, "uint4", torch.float16), (9, 1024, 640, "uint8b128", torch.float16),
                                      (3, 4096, 320, "uint4b8", torch.bfloat16)):  # fmt: skip
            out, want, _ = _case(m, k, n, wname, 128, dtype, True)
            assert_gemm_close(out, want, f"ring_kb={ring_kb} ctas={ctas}")
    finally:
        lib.cb200_set_stream_tuning(0, 0)


def _chain(pdl: bool, graph: bool):
    """A decode-token-like chain: GEMV i+1 consumes the output of GEMV i (so x is NOT static), 4 different layers."""
    lib = _lib.load()
    dev = torch.device("cuda")
    st = quant_utils.scalar_types.uint4
    dims = [(1024, 1536), (1536, 1024), (1024, 4096), (4096, 1024)] * 6
    layers = [quant_utils.random_packed_weights(k, n, st, 128, torch.float16, dev, seed=50 + i) for i, (k, n) in enumerate(dims)]
    x0 = (torch.randn((2, 1024), device=dev) * 0.5).to(torch.float16)
    acts = [x0] + [torch.empty((2, n), device=dev, dtype=torch.float16) for _, n in dims]

    def run():
        for i, (wq, ws, wzp) in enumerate(layers):
            mixed_precision_gemm(acts[i], wq, ws, wzp, 4, 0, 128, out=acts[i + 1], variant=STREAM)

    lib.cb200_set_pdl(0)
    run()
    torch.cuda.synchronize()
    want = [a.clone() for a in acts]
    try:
        lib.cb200_set_pdl(1 if pdl else 0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            if graph:
                run()
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    run()
                for _ in range(5):
                    for a in acts[1:]:
                        a.zero_()
                    g.replay()
            else:
                for _ in range(5):
                    for a in acts[1:]:
                        a.zero_()
                    run()
        torch.cuda.synchronize()
    finally:
        lib.cb200_set_pdl(0)
    for i, (a, w) in enumerate(zip(acts, want)):
        assert torch.equal(bits(a), bits(w)), f"activation {i} differs under pdl={pdl} graph={graph}"


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("pdl", [False, True], ids=["plain", "pdl"])
def test_dependent_chain_is_bit_identical_under_overlap(pdl, graph):
    _chain(pdl, graph)


def test_push_epilogue_writes_every_peer_buffer():
    """cb200_mpgemm_push: the last-arriver's vector stores also go to every peer pointer (here: buffers on the same GPU)."""
    p = oracle.make_problem(3, 1024, 448, "uint4", 128, torch.float16, True, seed=2)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    peers = [torch.zeros((3, 448), device="cuda", dtype=torch.float16) for _ in range(7)]
    out = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=STREAM, peer_out_ptrs=[t.data_ptr() for t in peers])
    torch.cuda.synchronize()
    for t in peers:
        assert torch.equal(bits(t), bits(out))
    assert_gemm_close(out, oracle.mixed_precision_gemm_ref(p.x, p.w_q_packed, p.w_s, p.w_zp, 4, 0, 128))


def test_chain_waits_inside_the_kernel_for_its_activations():
    """cb200_mpgemm_chain: x is a gather buffer "another GPU" fills (here: a second stream of the same GPU).  The GEMV is enqueued
    FIRST, its consumers spin on the completion counters while the ring fills; the writer then stores x and raises the counters
    (cb200_push with flags, the same counters a pushing GEMV raises).  The result must be computed from the NEW x, the wait's own
    count advances by one per call, and the pair replays."""
    import ctypes

    lib = _lib.load()
    p = oracle.make_problem(2, 2048, 512, "uint4", 128, torch.float16, True, seed=5)
    x, wq, ws, wzp = to_cuda(p.x, p.w_q_packed, p.w_s, p.w_zp)
    want = mixed_precision_gemm(x, wq, ws, wzp, 4, 0, 128, variant=STREAM)
    xbuf = torch.zeros_like(x)  # the "gather buffer": starts with stale contents
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")  # [0..1]: counters of two "ranks"; [8]: the chained kernel's own count
    out = torch.empty_like(want)
    writer, main = torch.cuda.Stream(), torch.cuda.Stream()
    # Every kernel the writer will launch is loaded BEFORE a GEMV spins: loading a module (CUDA lazy loading) while a kernel that
    # cannot finish is resident ends in a launch failure (measured: profiles/r02_chain_debug.log).
    with torch.cuda.stream(writer):
        torch.cuda._sleep(1000)
        scratch = torch.zeros_like(x)
        _lib.raise_for_status(lib.cb200_push(x.data_ptr(), x.numel() * 2, (ctypes.c_void_p * 1)(scratch.data_ptr()), 1, None, None,
                                             writer.cuda_stream))  # fmt: skip
    torch.cuda.synchronize()
    for rep in range(1, 4):
        xbuf.zero_()
        out.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(main):
            mixed_precision_gemm(xbuf, wq, ws, wzp, 4, 0, 128, out=out, x_ready=(flags.data_ptr(), 2, flags.data_ptr() + 32))
        assert lib.cb200_last_variant() == STREAM  # AUTO picks the kernel that can wait
        with torch.cuda.stream(writer):
            torch.cuda._sleep(2_000_000)  # the GEMV is resident and spinning by now
            peer = (ctypes.c_void_p * 1)(xbuf.data_ptr())
            pflag = (ctypes.c_void_p * 1)(flags.data_ptr() + 4)
            _lib.raise_for_status(lib.cb200_push(x.data_ptr(), x.numel() * 2, peer, 1, pflag, flags.data_ptr(), writer.cuda_stream))
        torch.cuda.synchronize()
        assert flags[:2].tolist() == [rep, rep] and int(flags[8]) == rep
        assert torch.equal(bits(out), bits(want))
    with pytest.raises(NotImplementedError):  # only the streaming kernel waits
        mixed_precision_gemm(xbuf, wq, ws, wzp, 4, 0, 128, out=out, variant=_lib.DECODE_MMA, x_ready=(flags.data_ptr(), 2, flags.data_ptr() + 32))
    with pytest.raises(ValueError):
        mixed_precision_gemm(xbuf, wq, ws, wzp, 4, 0, 128, out=out, x_ready=(flags.data_ptr(), 40, flags.data_ptr() + 32))


def test_push_kernel_copies_to_every_peer_and_counts():
    """cb200_push: the gather without the GEMM (side-stream form of the fused push)."""
    import ctypes

    lib = _lib.load()
    for nbytes, signalled in ((57344, True), (16, True), (3 << 20, False)):
        src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda")
        peers = [torch.zeros_like(src) for _ in range(7)]
        flags = torch.zeros(8, dtype=torch.int32, device="cuda")
        parr = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in peers])
        farr = (ctypes.c_void_p * 7)(*[flags.data_ptr() + 4 * (i + 1) for i in range(7)]) if signalled else None
        stream = torch.cuda.current_stream().cuda_stream
        for rep in range(2):
            _lib.raise_for_status(lib.cb200_push(src.data_ptr(), nbytes, parr, 7, farr, flags.data_ptr() if signalled else None, stream))
        torch.cuda.synchronize()
        for t in peers:
            assert torch.equal(t, src)
        assert flags.tolist() == ([2] * 8 if signalled else [0] * 8)
    with pytest.raises(ValueError):
        _lib.raise_for_status(lib.cb200_push(src.data_ptr(), 24, parr, 7, None, None, stream))  # not a multiple of 16
    with pytest.raises(NotImplementedError):
        _lib.raise_for_status(lib.cb200_push(src.data_ptr(), 3 << 20, parr, 7, farr or parr, flags.data_ptr(), stream))  # signalled: <= 1 MiB
    dst = torch.zeros_like(src)
    _lib.raise_for_status(lib.cb200_peer_copy(dst.data_ptr(), src.data_ptr(), src.numel(), stream))
    torch.cuda.synchronize()
    assert torch.equal(dst, src)


def test_dequant_still_bit_exact_next_to_it():
    p = oracle.make_problem(1, 1024, 256, "uint4", 128, torch.float16, True, seed=1)
    wq, ws, wzp = to_cuda(p.w_q_packed, p.w_s, p.w_zp)
    assert torch.equal(bits(dequantize_weights(wq, ws, wzp, 4, 0, 128)), bits(p.w_ref))
