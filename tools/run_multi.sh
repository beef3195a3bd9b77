#!/bin/bash
# usage: tools/run_multi.sh N   (on a box with N GPUs): multi-GPU parity check, then the weak-scaling bench at N
set -u
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/multi_parity_n$N.log 2>&1; echo "parity rc=$?"; grep "PASS\|FAIL\|unavailable" gpurun_out/multi_parity_n$N.log | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 500 --warmup 10 --per-step > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_n$N.err | head -3; tail -1 gpurun_out/bench_n$N.json | cut -c1-900
