set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
FL_TRACE=1 python tools/exp_frontend.py > gpurun_out/exp_fe.log 2>&1
grep "FE job\|stage us" gpurun_out/exp_fe.log
python tools/exp_staged.py > gpurun_out/exp_staged5.log 2>&1; tail -1 gpurun_out/exp_staged5.log
python bench.py --steps 300 --warmup 10 --no-cpu --no-icp > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; head -c 500 gpurun_out/bench_quick.json; echo; tail -c 300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
