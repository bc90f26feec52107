timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; tail -c 300 gpurun_out/bench_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print(d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['per_sample_ms'], d['other_schedules_per_gpu'])
PY
python profiles/configs_bench.py --json gpurun_out/configs_ab.json
