#!/bin/bash
# latency: parity test on the default lib, batch-1 timings (plain) for P1 / P0, then the phase profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "latency_configuration" > gpurun_out/r02_lat_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_lat_test.log
for P in P1 P0; do for m in 4 3; do TFHE_B200_LATENCY_CFG=$m timeout 300 python tools/prof_run.py --preset $P --batch 1 --steps 3 --warmup 1 --check --tag mode$m; done; done 2>&1 | tee gpurun_out/r02_lat_now.txt
LATMODE=4 TFHE_B200_LAT_PROF=1 timeout 300 python tools/latprof_run.py 2>&1 | grep "cycles per step" | awk 'NR%4==1' | tee gpurun_out/r02_lat_phases_now.txt
