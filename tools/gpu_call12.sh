#!/bin/bash
# round 2 final, one GPU: test suite, bench (both arms), launch list, ncu --set full of the blind rotation for P1 / P0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_v16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_v16.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v16_1gpu.json 2> gpurun_out/r02_bench_v16_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_v16_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_v16_reference_arm.json 2>> gpurun_out/r02_bench_v16_1gpu.err
M="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
for cfg in ; do
  set -- $cfg
  timeout 300 python tools/prof_run.py --preset $1 --batch $2 --steps 2 --warmup 1 --check > gpurun_out/r02_prof_plain_$1_v16.log 2>&1 &&
  timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:pbs_fft_kernel -s 1 -c 1 -f -o gpurun_out/r02_v16_$1 \
      python tools/prof_run.py --preset $1 --batch $2 --steps 1 --warmup 1 > gpurun_out/r02_ncu_$1_v16.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_v16.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_bench_v16.log 2>&1
tail -3 gpurun_out/r02_gputest_v16.log; tail -2 gpurun_out/r02_bench_v16_1gpu.err; cat gpurun_out/r02_prof_plain_P*_v16.log
