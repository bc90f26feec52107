# round 2, third GPU call: the pair-shared second Philox call (new sampler specification) — parity, then rates
timeout 1500 python -m pytest tests -m gpu -x -q -k "not statistics and not dropin and not cluster" > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2c.log
python bench.py --no-cpu-baseline > gpurun_out/bench_r2c_C4.json 2> gpurun_out/bench_r2c_C4.err; tail -c 600 gpurun_out/bench_r2c_C4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c_C4.json'))
print('C4', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel_ms'], d['roofline']['per_sample_ms'], d['other_schedules_per_gpu'], d['roofline']['compute_bound']['frac'])
PY
python profiles/configs_bench.py --json gpurun_out/configs_r2c.json
