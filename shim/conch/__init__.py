"""Drop-in shim: the reference's import paths for the mixed-precision GEMM hot path, served by conch_b200.

Put this directory in front of PYTHONPATH and the reference's own callers and tests
(tests/mixed_precision_gemm_test.py, benchmarks/mixed_precision_gemm_benchmark.py) run on the B200-native kernels
without a source change.  Only the modules that path imports exist here; everything else of the
reference is out of scope."""

import pathlib as _pathlib

# Modules this shim does not provide (the reference's own PyTorch oracles under conch.reference, conch.envs) are looked up in an
# installed copy of the reference when there is one next to the repo (baseline/_ref, the bench's reference arm): the hot path's
# modules -- ops, kernels, platforms, third_party -- always come from the shim, which is first on the package path.
_ref = _pathlib.Path(__file__).resolve().parents[2] / "baseline" / "_ref" / "conch"
if _ref.is_dir():
    __path__.append(str(_ref))
