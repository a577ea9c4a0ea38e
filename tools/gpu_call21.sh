#!/bin/bash
# BMMP with the tensor-memory tail: BMMP tests, then the BMMP bench with both exchange modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_exactness.py tests/test_gpu_fullsize.py -x -q -m gpu -k "bmmp" > gpurun_out/r02_bmmp_test.log 2>&1
echo "test rc=$?"; tail -5 gpurun_out/r02_bmmp_test.log
for rep in 1 2; do for tm in 1 0; do TFHE_B200_FFT_TMEM=$tm timeout 300 python tools/bmmp_bench.py 2>&1 | tail -1 | sed "s/^/tmem$tm /"; done; done | tee gpurun_out/r02_bmmp_tail_ab.txt
