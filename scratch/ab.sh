python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for lib in libmcrg_b200.so libmcrg_mb3.so libmcrg_old.so; do
  echo "== $lib"
  MCRG_LIB=$PWD/mcrg_b200/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['per_sample_ms'])"
done
