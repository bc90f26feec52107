timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_benched_path.py -x -q -k "pyramid or run_accumulators or graph or sharded or supplied or golden or full_size" > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2g.log
python profiles/configs_bench.py 2>&1 | tail -8
