# round 2, first GPU call: GPU tests incl. the benched-path parity tests, bench, host probe, sanitizer on HEAD's kernels
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2a.log
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 400 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print(d['value'], d['e2e'], d['roofline']['kernel_ms'], d['roofline']['per_sample_ms'], d['other_schedules_per_gpu'], d.get('cpu_baseline'))
PY
python profiles/host_probe.py --json gpurun_out/host_probe_r2a.json 2>&1 | tail -12
lscpu > gpurun_out/lscpu_r2a.txt; nvidia-smi topo -m > gpurun_out/topo_r2a.txt 2>&1
( timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize.py all ) > gpurun_out/sanitizer_memcheck_r2.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_r2.txt
( timeout 1200 compute-sanitizer --tool racecheck python profiles/sanitize.py sweep ) > gpurun_out/sanitizer_racecheck_r2.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_r2.txt
python profiles/configs_bench.py --json gpurun_out/configs_r2a.json
