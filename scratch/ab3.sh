python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['per_sample_ms'])"; }
echo "prefetch R=64"; run
echo "prefetch R=32"; run --strip-rows 32
echo "prefetch R=128"; run --strip-rows 128
echo "no prefetch R=64"; MCRG_PREFETCH=0 run
