# round 2, second GPU call: full GPU tests after the M^4 / ragged-strip / upload / drop-in changes, strip-height scan, every config
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2b.log
python profiles/strip_scan.py --json gpurun_out/strip_scan_r2b.json 2>&1 | tee gpurun_out/strip_scan_r2b.txt
python bench.py > gpurun_out/bench_r2b_C4.json 2> gpurun_out/bench_r2b_C4.err; tail -c 600 gpurun_out/bench_r2b_C4.err
for c in C1 C2 C3 C5; do python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/bench_r2b_$c.json 2> gpurun_out/bench_r2b_$c.err; tail -c 300 gpurun_out/bench_r2b_$c.err; done
python - <<'PY'
import json
for c in ("C4","C1","C2","C3","C5"):
    try:
        d=json.load(open(f'gpurun_out/bench_r2b_{c}.json'))
        print(c, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['path'], {k:round(v,1) for k,v in d['e2e']['variants'].items()}, 'roof', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e']['host_roofline'].items() if k!='what'}, d['roofline']['kernel_ms'], d['roofline']['per_sample_ms'], d['other_schedules_per_gpu'], d['config']['kernel_path'])
    except Exception as e:
        print(c, 'failed', e)
PY
