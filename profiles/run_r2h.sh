timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py -x -q -k "sweeps or resident or run_accumulators or checkpoint or sharded or thermo or lambda or stat" > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2h.log
python profiles/configs_bench.py --only "C1" 2>&1 | tail -1
python profiles/configs_bench.py --only "C2" 2>&1 | tail -1
