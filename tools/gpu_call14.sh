#!/bin/bash
# P1 tensor-memory tail variants: parity test (default lib), A/B of variants/lib_*.so, then ncu of the default lib
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exactness.py -x -q -m gpu -k "tensor_memory" > gpurun_out/r02_tmem_test.log 2>&1
echo "test rc=$?"; tail -3 gpurun_out/r02_tmem_test.log
rm -f gpurun_out/r02_tail_variants.txt
for rep in 1 2; do bash tools/ab_run.sh gpurun_out/r02_tail_variants.txt P1 4096 "$@"; done
TFHE_B200_FFT_TMEM=0 timeout 300 python tools/prof_run.py --preset P1 --batch 4096 --steps 3 --warmup 1 --check --tag smem >> gpurun_out/r02_tail_variants.txt
cat gpurun_out/r02_tail_variants.txt; tail -3 gpurun_out/r02_tail_variants.txt.err 2>/dev/null

exit 0
M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,lts__t_bytes.sum"
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:pbs_fft_kernel -s 1 -c 1 -f -o gpurun_out/r02_v14_P1 \
    python tools/prof_run.py --preset P1 --batch 4096 --steps 1 --warmup 1 > gpurun_out/r02_ncu_P1_v14.log 2>&1
tail -2 gpurun_out/r02_ncu_P1_v14.log
