#!/bin/bash
# A/B of variants on P1 (4096) and P2 (2368) after the tensor-memory parity test on the default lib
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_tmem_test.log
rm -f gpurun_out/r02_variants_p1p2.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_variants_p1p2.txt P1 4096 "$@"; bash tools/ab_run.sh gpurun_out/r02_variants_p1p2.txt P2 2368 "$@"; done
cat gpurun_out/r02_variants_p1p2.txt; tail -3 gpurun_out/r02_variants_p1p2.txt.err 2>/dev/null
