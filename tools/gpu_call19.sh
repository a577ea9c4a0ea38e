#!/bin/bash
# round 2 closing run, one GPU: whole test suite, smoke, bench (both arms), ncu --set full of the latency (cluster) kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_v16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_v16.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_v16.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke_v16.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v16_1gpu.json 2> gpurun_out/r02_bench_v16_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_v16_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_v16_reference_arm.json 2>> gpurun_out/r02_bench_v16_1gpu.err
M="sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,lts__t_bytes.sum"
timeout 300 python tools/prof_run.py --preset P1 --batch 1 --steps 2 --warmup 1 --check > gpurun_out/r02_prof_plain_lat_v16.log 2>&1 &&
timeout 600 ncu --set full --metrics $M --clock-control none --import-source on -k regex:cluster_split -s 1 -c 1 -f -o gpurun_out/r02_v16_cluster_split_P1 \
    python tools/prof_run.py --preset P1 --batch 1 --steps 1 --warmup 1 > gpurun_out/r02_ncu_lat_v16.log 2>&1
tail -3 gpurun_out/r02_gputest_v16.log; tail -2 gpurun_out/r02_smoke_v16.log; tail -2 gpurun_out/r02_bench_v16_1gpu.err; cat gpurun_out/r02_prof_plain_lat_v16.log; tail -1 gpurun_out/r02_ncu_lat_v16.log
