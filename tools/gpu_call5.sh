#!/bin/bash
# ncu captures of the round-2 kernels: ks_tcgen05_kernel (P1 batch 4096) and pbs_fft_cluster_split_kernel (P1 single PBS)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
timeout 300 python tools/prof_run.py --preset P1 --batch 4096 --steps 1 --warmup 1 --check > gpurun_out/r02_prof_plain_ks.log 2>&1 &&
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:ks_tcgen05 -s 1 -c 1 -f -o gpurun_out/r02_ks_tcgen05 \
    python tools/prof_run.py --preset P1 --batch 4096 --steps 1 --warmup 1 > gpurun_out/r02_ncu_ks.log 2>&1
timeout 300 python tools/prof_run.py --preset P1 --batch 1 --steps 1 --warmup 1 --check > gpurun_out/r02_prof_plain_lat.log 2>&1 &&
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:cluster_split -s 1 -c 1 -f -o gpurun_out/r02_cluster_split_P1 \
    python tools/prof_run.py --preset P1 --batch 1 --steps 1 --warmup 1 > gpurun_out/r02_ncu_lat.log 2>&1
tail -2 gpurun_out/r02_ncu_ks.log gpurun_out/r02_ncu_lat.log; cat gpurun_out/r02_prof_plain_ks.log gpurun_out/r02_prof_plain_lat.log
