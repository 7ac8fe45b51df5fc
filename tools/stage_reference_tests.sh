#!/usr/bin/env bash
# Stage the reference's OWN test of the hot path next to the installed reference (baseline/, git-ignored, shipped to the GPU box
# by gpurun) so that it can run UNMODIFIED against the conch -> conch_b200 shim:
#     tools/stage_reference_tests.sh && gpurun -- 'PYTHONPATH=shim:. python -m pytest baseline/_ref_tests -q'
# Nothing from /root/reference is copied into the tracked tree.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
mkdir -p "$ROOT/baseline/_ref_tests"
for t in mixed_precision_gemm_test.py scaled_gemm_test.py int8_quant_kernels_test.py fp8_quant_kernels_test.py; do
  cp "$REF/tests/$t" "$ROOT/baseline/_ref_tests/"
done
echo "staged $(ls "$ROOT/baseline/_ref_tests")"
