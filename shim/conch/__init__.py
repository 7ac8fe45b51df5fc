"""Drop-in shim: the reference's import paths for the mixed-precision GEMM hot path, served by conch_b200.

Put this directory in front of PYTHONPATH and the reference's own callers and tests
(tests/mixed_precision_gemm_test.py, benchmarks/mixed_precision_gemm_benchmark.py) run on the B200-native kernels
without a source change.  Only the modules that path imports exist here; everything else of the
reference is out of scope."""
