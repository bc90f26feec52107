nproc; python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upload or roundtrip or sweeps_match" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b5.err | tail -1 > gpurun_out/bench_r1h.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r1h.json')); print(d['value'], d['ms_per_step'], d['e2e'])"; tail -2 gpurun_out/b5.err
