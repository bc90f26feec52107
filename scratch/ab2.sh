python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_sample_ms'])"
done
