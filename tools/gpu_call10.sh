#!/bin/bash
# A/B of variants on P0 and P1 after the exactness tests on the default lib
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_exact_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_exact_test.log
rm -f gpurun_out/r02_variants_p0p1.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_variants_p0p1.txt P0 4096 "$@"; bash tools/ab_run.sh gpurun_out/r02_variants_p0p1.txt P1 4096 "$@"; done
cat gpurun_out/r02_variants_p0p1.txt; tail -3 gpurun_out/r02_variants_p0p1.txt.err 2>/dev/null
