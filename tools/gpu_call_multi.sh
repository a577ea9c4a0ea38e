#!/bin/bash
# multi-GPU call: N-GPU == 1-GPU tests (NCCL ranks + tfhe_mgpu C ABI) and the sharded bench line at N GPUs
# usage: bash tools/gpu_call_multi.sh N [steps] [skip_tests]
cd "$(dirname "$0")/.."
N=$1; STEPS=${2:-5}; SKIPT=${3:-0}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_smi_${N}gpu.txt 2>&1
if [ "$SKIPT" = "0" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q -s > gpurun_out/r02_gputest_multigpu_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_multigpu_${N}gpu.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 \
  > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/r02_bench_${N}gpu.err
tail -4 gpurun_out/r02_gputest_multigpu_${N}gpu.log 2>/dev/null; tail -c 1500 gpurun_out/r02_bench_${N}gpu.err; head -c 600 gpurun_out/r02_bench_${N}gpu.json
