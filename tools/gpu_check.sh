#!/bin/bash
# Standard GPU-box sequence: parity tests, bench (both arms), ncu launch list + full captures of the two dominant kernels
# (profiling runs never produce bench values).  usage: tools/gpu_check.sh [tag]
set -u
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 500 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu --no-icp > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_similarity_staged -s 4 -c 1 -f -o gpurun_out/prof_sim_$TAG \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_full.log 2>&1; echo "ncu full sim rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_front_end_wave -s 4 -c 1 -f -o gpurun_out/prof_fe_$TAG \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_full_fe.log 2>&1; echo "ncu full fe rc=$?"
# the ICP loop kernel (C3 batch inside bench.py's icp leg; the Recognition leg launches it with 5 hypotheses)
ncu --set full --clock-control none --import-source on -k regex:k_icp_fused -s 1 -c 1 -f -o gpurun_out/prof_icp_$TAG \
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_icp.log 2>&1; echo "ncu full icp rc=$?"
