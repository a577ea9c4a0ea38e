#!/bin/bash
# P2 tail variants: parity test (default lib), A/B of variants/lib_*.so at batch 2368 + the shared-memory kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_tmem_test.log
rm -f gpurun_out/r02_tail16_variants.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_tail16_variants.txt P2 2368 "$@"; done
TFHE_B200_FFT_TMEM=0 timeout 300 python tools/prof_run.py --preset P2 --batch 2368 --steps 2 --warmup 1 --check --tag smem >> gpurun_out/r02_tail16_variants.txt
cat gpurun_out/r02_tail16_variants.txt; tail -3 gpurun_out/r02_tail16_variants.txt.err 2>/dev/null
