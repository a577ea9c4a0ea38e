#!/bin/bash
# latency A/B: latency-mode parity test on the default lib, then batch-1 PBS of P1 and P0 with variants/lib_*.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "latency_configuration" > gpurun_out/r02_lat_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_lat_test.log
rm -f gpurun_out/r02_lat_variants.txt
for rep in 1 2; do for P in P1 P0; do bash tools/ab_run.sh gpurun_out/r02_lat_variants.txt $P 1 "$@"; done; done
cat gpurun_out/r02_lat_variants.txt; tail -3 gpurun_out/r02_lat_variants.txt.err 2>/dev/null
