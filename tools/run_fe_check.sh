set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in "" "FL_PDL_REFINE=1" "FL_NO_PDL_SIM=1"; do
  env $v python bench.py --steps 300 --warmup 10 --no-cpu --no-icp > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "variant [$v]"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_quick.json'))
print('ms_per_step', b['ms_per_step'], 'p50', b['ms_per_step_p50'], 'e2e fps', b['e2e']['frames_per_s'], 'stages', b['stage_ms'], 'frac', b['roofline']['frac'])
PY
done
