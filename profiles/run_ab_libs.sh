# A/B of library builds under gpurun_ab/lib*.so (same sources, different compile-time choices): rates per config
for lib in gpurun_ab/lib*.so; do
  echo "=== $lib"
  MCRG_LIB=$PWD/$lib python profiles/configs_bench.py --only "x 40" 2>&1 | tail -1
  MCRG_LIB=$PWD/$lib python profiles/configs_bench.py --only "C3" 2>&1 | tail -1
  MCRG_LIB=$PWD/$lib python profiles/configs_bench.py --only "C5" 2>&1 | tail -1
  MCRG_LIB=$PWD/$lib python profiles/configs_bench.py --only "C2" 2>&1 | tail -1
done
