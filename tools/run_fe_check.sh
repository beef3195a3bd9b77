set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python tools/exp_frontend.py > gpurun_out/exp_fe.log 2>&1; cat gpurun_out/exp_fe.log
python bench.py --steps 300 --warmup 10 --no-cpu --no-icp > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fe2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-icp > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
