python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['per_sample_ms'])"; }
echo "pair mb3 no prefetch"; MCRG_PREFETCH=0 run
echo "pair mb3 prefetch"; run
echo "pair mb2 no prefetch"; MCRG_LIB=$PWD/mcrg_b200/libmcrg_mb2.so MCRG_PREFETCH=0 run
echo "pair mb2 prefetch"; MCRG_LIB=$PWD/mcrg_b200/libmcrg_mb2.so run
echo "pair mb2 no prefetch R=128"; MCRG_LIB=$PWD/mcrg_b200/libmcrg_mb2.so MCRG_PREFETCH=0 run --strip-rows 128
