#!/bin/bash
# ncu capture of the tensor-memory-exchange kernel (P0, batch 4096), the shared-memory-exchange kernel timed beside it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,lts__t_bytes.sum"
timeout 300 python tools/prof_run.py --preset P0 --batch 4096 --steps 2 --warmup 1 --check --tag tmem1 > gpurun_out/r02_tmem_plain.log 2>&1 &&
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:pbs_fft_kernel -s 1 -c 1 -f -o gpurun_out/r02_v12_P0 \
    python tools/prof_run.py --preset P0 --batch 4096 --steps 1 --warmup 1 > gpurun_out/r02_ncu_tmem.log 2>&1
TFHE_B200_FFT_TMEM=0 timeout 300 python tools/prof_run.py --preset P0 --batch 4096 --steps 2 --warmup 1 --check --tag tmem0 >> gpurun_out/r02_tmem_plain.log 2>&1
tail -2 gpurun_out/r02_ncu_tmem.log; cat gpurun_out/r02_tmem_plain.log
