#!/bin/bash
# A/B of tensor-memory-exchange variants (P0): parity test on the default lib, then timing of variants/lib_*.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -5 gpurun_out/r02_tmem_test.log
rm -f gpurun_out/r02_tmem_variants.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_tmem_variants.txt P0 4096 "$@"; done
cat gpurun_out/r02_tmem_variants.txt; tail -3 gpurun_out/r02_tmem_variants.txt.err 2>/dev/null
