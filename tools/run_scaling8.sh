#!/bin/bash
# On an 8-GPU box: multi-GPU parity at world 8, the C2 weak-scaling point and the C4 / C5 strong-scaling points at N = 8,
# then the multi-device C++ / group tests.  Results land in gpurun_out/.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nproc > gpurun_out/box_cores.txt
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/multi_parity_n8.log 2>&1; echo "parity8 rc=$?"; grep "PARITY\|FAIL\|unavailable" gpurun_out/multi_parity_n8.log | tail -4
timeout 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "C2 n8 rc=$?"; tail -1 gpurun_out/bench_n8.json | cut -c1-200
timeout 400 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --config C4 --steps 200 --warmup 3 > gpurun_out/pipe_c4_n8.json 2> gpurun_out/pipe_c4_n8.err; echo "C4 n8 rc=$?"; tail -1 gpurun_out/pipe_c4_n8.json | cut -c1-200
timeout 500 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --config C5 --steps 100 --warmup 3 > gpurun_out/pipe_c5_n8.json 2> gpurun_out/pipe_c5_n8.err; echo "C5 n8 rc=$?"; tail -1 gpurun_out/pipe_c5_n8.json | cut -c1-200
timeout 400 python -m pytest tests/test_cpp_shim.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi8.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi8.log
