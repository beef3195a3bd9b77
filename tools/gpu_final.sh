#!/bin/bash
# full GPU suite, smoke, bench (both arms), launch list, one full capture of the similarity kernel.  usage: tools/gpu_final.sh [tag]
set -u
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_similarity_staged -s 12 -c 1 -f -o gpurun_out/prof_sim_$TAG \
  python bench.py --steps 8 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_full_sim.log 2>&1; echo "ncu full sim rc=$?"
