#!/bin/bash
# P2 tensor-memory tail: parity test, then A/B against the shared-memory kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -25 gpurun_out/r02_tmem_test.log
for rep in 1 2; do for tm in 1 0; do
  TFHE_B200_FFT_TMEM=$tm timeout 300 python tools/prof_run.py --preset P2 --batch 2368 --steps 2 --warmup 1 --check --tag tmem$tm
done; done 2>&1 | tee gpurun_out/r02_tail16_ab.txt
