set -x
python profiles/configs_bench.py --only C2 2>&1 | tee gpurun_out/ab_c2_new.txt
MCRG_RESIDENT_THREADS=96 python profiles/configs_bench.py --only C2 2>&1 | tee gpurun_out/ab_c2_old.txt
MCRG_RESIDENT_THREADS=64 python profiles/configs_bench.py --only C2 2>&1 | tee gpurun_out/ab_c2_64.txt
python profiles/configs_bench.py --only "L=128" 2>&1 | tee gpurun_out/ab_128_new.txt
MCRG_RESIDENT_THREADS=256 python profiles/configs_bench.py --only "L=128" 2>&1 | tee gpurun_out/ab_128_old.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 600 gpurun_out/bench_r1b.err; cut -c1-400 gpurun_out/bench_r1b.json
