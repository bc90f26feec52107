# Run on the B200 (gpurun): GPU tests, the bench (every config), the reference arm, the ncu launch list, one ncu --set full capture of
# the measuring sweep kernel and the per-configuration rates; outputs land in gpurun_out/, `python profiles/summarize.py r2_final`
# condenses them into profiles/.
TAG=${TAG:-r2_final}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
for c in C1 C2 C3 C5; do python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --samples 32 --e2e-steps 2 --no-cpu-baseline > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 40 -c 1 -o gpurun_out/prof_sweep_$TAG -f python bench.py --steps 1 --warmup 1 --samples 16 --graphs 0 --e2e-steps 2 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_sweep_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv
ncu -i gpurun_out/prof_sweep_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$TAG.csv 2>&1
python profiles/configs_bench.py --json gpurun_out/configs_$TAG.json > gpurun_out/configs_$TAG.txt 2>&1
python profiles/cluster_bench.py > gpurun_out/cluster_bench_$TAG.txt 2>&1
python - <<PY
import json
for c in ("","_C1","_C2","_C3","_C5"):
    try:
        d=json.load(open(f'gpurun_out/bench_${TAG}{c}.json'))
        print(d['config']['name'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['path'], 'roofline frac', round(d['roofline']['frac'],4), 'whole step', round(d['roofline']['whole_step_frac'],4), 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['per_sample_ms'], d['other_schedules_per_gpu'], d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(c,'failed',e)
print(open('gpurun_out/bench_ref_${TAG}.json').read()[:300])
PY
cat gpurun_out/configs_$TAG.txt
