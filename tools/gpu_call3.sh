#!/bin/bash
# round 2, GPU call 3: latency kernel (all teams on one ciphertext): parity tests, latency table, margins of the adversarial tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exactness.py -m gpu -x -q -s > gpurun_out/r02_gputest_exactness.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_exactness.log
timeout 600 python tools/latency_run.py > gpurun_out/r02_latency.json 2> gpurun_out/r02_latency.err
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
grep -E "margin|passed|failed|rc=" gpurun_out/r02_gputest_exactness.log | tail -20; cat gpurun_out/r02_latency.json; tail -5 gpurun_out/r02_latency.err; tail -3 gpurun_out/r02_gputest.log
