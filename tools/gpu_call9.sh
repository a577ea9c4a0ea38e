#!/bin/bash
# full GPU suite + default bench + ncu of the P0 tensor-memory kernel at batch 4096
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_v12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_v12.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v12_1gpu.json 2> gpurun_out/r02_bench_v12_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_v12_1gpu.err
FP64M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
timeout 300 python tools/prof_run.py --preset P0 --batch 4096 --steps 2 --warmup 1 --check > gpurun_out/r02_prof_plain_P0_v12.log 2>&1 &&
timeout 900 ncu --set full --metrics $FP64M --clock-control none --import-source on -k regex:pbs_fft_kernel -s 1 -c 1 -f -o gpurun_out/r02_v12_P0 \
    python tools/prof_run.py --preset P0 --batch 4096 --steps 1 --warmup 1 > gpurun_out/r02_ncu_P0_v12.log 2>&1
tail -3 gpurun_out/r02_gputest_v12.log; tail -2 gpurun_out/r02_bench_v12_1gpu.err; cat gpurun_out/r02_prof_plain_P0_v12.log; tail -2 gpurun_out/r02_ncu_P0_v12.log
