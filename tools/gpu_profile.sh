#!/bin/bash
# bench (both arms) + ncu launch list + full captures of the front end and the ICP kernel.  usage: tools/gpu_profile.sh [tag]
set -u
TAG=${1:-cur}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_front_end_wave -s 12 -c 1 -f -o gpurun_out/prof_fe_$TAG \
  python bench.py --steps 8 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_full_fe.log 2>&1; echo "ncu full fe rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_icp_fused -s 1 -c 1 -f -o gpurun_out/prof_icp_$TAG \
  python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_full_icp.log 2>&1; echo "ncu full icp rc=$?"
