# Run on the B200 (gpurun): the bench, the reference arm, the ncu launch list, one ncu --set full capture of the measuring
# sweep kernel and the per-configuration rates; outputs land in gpurun_out/, `python profiles/summarize.py r1_final` condenses them.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1_final.json 2>> gpurun_out/bench_r1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --samples 32 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 40 -c 1 -o gpurun_out/prof_sweep_r1_final -f python bench.py --steps 1 --warmup 1 --samples 16 --graphs 0 --no-cpu-baseline > gpurun_out/ncu_r1_final.log 2>&1
ncu -i gpurun_out/prof_sweep_r1_final.ncu-rep --page raw --csv > gpurun_out/raw_r1_final.csv
ncu -i gpurun_out/prof_sweep_r1_final.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_r1_final.csv 2>&1
python profiles/configs_bench.py --json gpurun_out/configs_r1_final.json > gpurun_out/configs_r1_final.txt 2>&1
python profiles/cluster_bench.py > gpurun_out/cluster_bench_r1_final.txt 2>&1
tail -c 1500 gpurun_out/bench_r1_final.json; cat gpurun_out/bench_ref_r1_final.json | cut -c1-300
