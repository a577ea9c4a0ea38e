#!/bin/bash
# A/B of variants on one preset/batch after the exactness tests on the default lib:  gpu_call11.sh <preset> <batch> <variants...>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
preset=$1; batch=$2; shift 2
timeout 1200 python -m pytest tests/test_gpu_exactness.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_exact_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_exact_test.log
rm -f gpurun_out/r02_variants_$preset.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_variants_$preset.txt $preset $batch "$@"; done
cat gpurun_out/r02_variants_$preset.txt; tail -3 gpurun_out/r02_variants_$preset.txt.err 2>/dev/null
