"""CPU: host-side logic -- the C-ABI library builds, loads and exports exactly what include/conch_b200.h declares
(no compute calls: there is no GPU here), metadata deduction / validation mirrors the reference op, the packaged
quantiser agrees with the oracle, and the op refuses to run without a B200 (no CPU fallback)."""

from __future__ import annotations

import ctypes
import pathlib
import re

import pytest
import torch

import mpgemm_oracle as oracle
from conch_b200 import _lib, quant_utils
from conch_b200.kernels.quantization.gemm import MixedPrecisionMatmulMetadata, WeightGroupMode
from conch_b200.ops.quantization.gemm import create_mixed_precision_metadata, mixed_precision_gemm

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _declared_symbols() -> list[str]:
    header = (ROOT / "include" / "conch_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z_0-9]+)\s*\(", header)))


def test_header_declares_what_the_binding_binds():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol(built_lib):
    assert built_lib.exists()
    lib = ctypes.CDLL(str(built_lib))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {built_lib.name}"
    typed = _lib.load()
    assert typed.cb200_version() == 100
    # workspace sizing is pure host arithmetic -- safe without a GPU
    assert typed.cb200_mpgemm_workspace_bytes(1, 4096, 4096, 4, 0) >= _lib.WORKSPACE_COUNTER_BYTES


def test_library_is_sm100a_only(built_lib):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def _cpu_problem(m=3, k=256, n=128, wtype="uint4", g=128, dtype=torch.float16, zp=True):
    return oracle.make_problem(m, k, n, wtype, g, dtype, zp, seed=1)


def test_metadata_matches_reference_deduction():
    p = _cpu_problem()
    md = create_mixed_precision_metadata(*p.op_args)
    assert isinstance(md, MixedPrecisionMatmulMetadata)
    assert (md.m_dim, md.k_dim, md.n_dim) == (3, 256, 128)
    assert md.elements_per_sample == 8 and md.unpack_mask == 15
    assert md.zero_is_scalar is False
    assert md.input_dtype == md.output_dtype == md.meta_dtype == torch.float16
    assert md.acc_dtype == torch.float32
    assert md.weight_group_mode is WeightGroupMode.SYMMETRIC_WITH_SHIFT
    md8 = create_mixed_precision_metadata(*_cpu_problem(wtype="uint8b128", zp=False).op_args, output_dtype=torch.float32)
    assert md8.elements_per_sample == 4 and md8.unpack_mask == 255 and md8.weight_bias == 128
    assert md8.weight_group_mode is WeightGroupMode.SYMMETRIC_NO_SHIFT
    assert md8.output_dtype == torch.float32


def test_group_size_minus_one_is_per_channel():
    p = _cpu_problem(g=-1)
    assert p.w_s.shape == (1, 128)
    md = create_mixed_precision_metadata(*p.op_args)
    assert md.group_size == 256


def test_scalar_zero_point_detected():
    p = _cpu_problem()
    args = list(p.op_args)
    args[3] = torch.tensor([[7]], dtype=torch.int32)
    assert create_mixed_precision_metadata(*args).zero_is_scalar is True
    args[3] = torch.tensor([7], dtype=torch.int32)  # rank-1 scalar tolerated unless strict
    assert create_mixed_precision_metadata(*args).zero_is_scalar is True
    with pytest.raises(ValueError):
        create_mixed_precision_metadata(*args, strict=True)


@pytest.mark.parametrize("strict", [True, False])
def test_validation_errors_match_reference_types(strict):
    p = _cpu_problem()
    x, wq, ws, wzp, bits, bias, g = p.op_args
    with pytest.raises(ValueError, match="dimensions of input tensor x"):
        create_mixed_precision_metadata(x[0], wq, ws, wzp, bits, bias, g, strict=strict)
    with pytest.raises(ValueError, match="Invalid datatype for packed weights"):
        create_mixed_precision_metadata(x, wq.to(torch.int64), ws, wzp, bits, bias, g, strict=strict)
    with pytest.raises(ValueError, match="Invalid w_s shape"):
        create_mixed_precision_metadata(x, wq, ws[:1], wzp, bits, bias, g, strict=strict)
    with pytest.raises(ValueError, match="Invalid w_zp shape"):
        create_mixed_precision_metadata(x, wq, ws, wzp[:, :64], bits, bias, g, strict=strict)
    with pytest.raises(NotImplementedError):
        create_mixed_precision_metadata(x, wq, ws, wzp, 2, bias, g, strict=strict)
    with pytest.raises(NotImplementedError):
        create_mixed_precision_metadata(x.float(), wq, ws, wzp, bits, bias, g, strict=strict)
    with pytest.raises(NotImplementedError):
        create_mixed_precision_metadata(x, wq, ws, wzp, bits, bias, g, acc_dtype=torch.float16, strict=strict)


def test_scaled_activations_only_rejected_in_strict_mode():
    p = _cpu_problem()
    with pytest.raises(NotImplementedError, match="Scaled activations"):
        create_mixed_precision_metadata(*p.op_args, scaled_activations=True, strict=True)
    create_mixed_precision_metadata(*p.op_args, scaled_activations=True, strict=False)  # ignored, like the reference


def test_no_cpu_fallback():
    p = _cpu_problem()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mixed_precision_gemm(*p.op_args)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "conch_b200").rglob("*.py"):
        text = path.read_text()
        assert "mpgemm_oracle" not in text and "import oracle" not in text and "from oracle" not in text, path


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("wname", ["uint4", "uint4b8", "uint8", "uint8b128"])
@pytest.mark.parametrize("zp", [True, False])
@pytest.mark.parametrize("g", [64, 128, -1])
def test_packaged_quantiser_equals_oracle(dtype, wname, zp, g):
    gen = torch.Generator().manual_seed(3)
    w = (10 * (torch.rand((256, 64), generator=gen) - 0.3)).to(dtype)
    st = getattr(quant_utils.scalar_types, wname)
    a = quant_utils.quantize_weights(w, st, g, zp)
    b = oracle.quantize_weights(w, oracle.WEIGHT_TYPES[wname], g, zp)
    for u, v in zip(a, b):
        assert (u is None and v is None) or torch.equal(u, v)
    assert torch.equal(quant_utils.pack_rows(a[1], st.size_bits), oracle.pack_rows(b[1], st.size_bits))
    assert (st.size_bits, st.bias) == (oracle.WEIGHT_TYPES[wname].bits, oracle.WEIGHT_TYPES[wname].bias)


def test_workspace_sizes_stay_small_for_the_llama_shapes():
    """Scratch is split-K partials only: tile-local for the tcgen05 path (never an [M, N] plane per split), a few fp32 rows
    per column block for the decode path.  Pure host arithmetic in the library -- runs without a GPU."""
    lib = _lib.load()
    mib = 1 << 20
    for m, k, n in [(1, 4096, 4096), (1, 4096, 28672), (8, 14336, 4096), (16, 4096, 28672), (64, 4096, 4096),
                    (512, 4096, 28672), (2048, 4096, 14336), (4096, 4096, 14336), (4096, 8192, 4096), (4096, 8192, 28672)]:
        for bits in (4, 8):
            need = lib.cb200_mpgemm_workspace_bytes(m, k, n, bits, _lib.AUTO)
            assert need == 0 or need >= _lib.WORKSPACE_COUNTER_BYTES
            assert need < 96 * mib, (m, k, n, bits, need)
