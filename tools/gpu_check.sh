#!/bin/bash
# Rebuild the product library, then run the GPU parity tests and the phase-timing script on a B200 (gpurun).
set -e
cd "$(dirname "$0")/.."
make -C alignasm_b200/csrc all 2>&1 | grep -E "error|Error" && exit 1
/usr/local/graft/bin/gpurun --timeout 900 -- "python -m pytest tests -m gpu -x -q 2>&1 | tail -6; timeout 600 python tools/first_gpu.py ${1:-} 2>&1 | tail -12"
