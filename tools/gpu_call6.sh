#!/bin/bash
# tensor-memory-exchange kernel (P0): parity, then A/B timing against the shared-memory exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -15 gpurun_out/r02_tmem_test.log
for tm in 1 0; do
  TFHE_B200_FFT_TMEM=$tm timeout 300 python tools/prof_run.py --preset P0 --batch 4096 --steps 3 --warmup 1 --check --tag tmem$tm
  TFHE_B200_FFT_TMEM=$tm timeout 300 python tools/prof_run.py --preset P0 --batch 592 --steps 3 --warmup 1 --check --tag tmem$tm
done 2>&1 | tee gpurun_out/r02_tmem_ab.txt
