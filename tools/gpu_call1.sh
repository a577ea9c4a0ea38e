#!/bin/bash
# round 2, GPU call 1: test suite, default bench line, launch list, ncu --set full of the blind rotation for P1 / P0 / P2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_1gpu.err
FP64M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
for cfg in "P1 4096" "P0 4096" "P2 2368"; do
  set -- $cfg
  timeout 300 python tools/prof_run.py --preset $1 --batch $2 --steps 1 --warmup 1 --check > gpurun_out/r02_prof_plain_$1.log 2>&1 &&
  timeout 900 ncu --set full --metrics $FP64M --clock-control none --import-source on -k regex:pbs_fft_kernel -s 1 -c 1 -f -o gpurun_out/r02_$1 \
      python tools/prof_run.py --preset $1 --batch $2 --steps 1 --warmup 1 > gpurun_out/r02_ncu_$1.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/r02_gputest.log
