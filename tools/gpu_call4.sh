#!/bin/bash
# full validation of the current tree on one B200: test suite, smoke, default bench line, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 python tools/latency_run.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err
tail -3 gpurun_out/r02_gputest.log; cat gpurun_out/r02_smoke.log; tail -c 400 gpurun_out/r02_bench_1gpu.err; head -c 300 gpurun_out/r02_bench_1gpu.json; echo; head -c 400 gpurun_out/r02_bench_reference.json
