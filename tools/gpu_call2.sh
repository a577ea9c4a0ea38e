#!/bin/bash
# round 2, GPU call 2: microbenchmark (duplicate-address LDS.128), latency configurations, test suite, bench line, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/bin/lds_dup > gpurun_out/r02_lds_dup.json 2>&1
timeout 600 python tools/latency_run.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_1gpu.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_bench_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
cat gpurun_out/r02_lds_dup.json; cat gpurun_out/r02_latency.json; tail -3 gpurun_out/r02_gputest.log; tail -c 600 gpurun_out/r02_bench_1gpu.err
