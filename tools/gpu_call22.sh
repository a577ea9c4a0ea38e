#!/bin/bash
# closing check of the committed state: whole GPU test suite, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final2_1gpu.json 2> gpurun_out/r02_bench_final2_1gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_final2_1gpu.err
tail -3 gpurun_out/r02_gputest_final.log; tail -2 gpurun_out/r02_smoke_final.log; tail -1 gpurun_out/r02_bench_final2_1gpu.err; head -c 400 gpurun_out/r02_bench_final2_1gpu.json
