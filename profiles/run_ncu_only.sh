TAG=r2_final
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --samples 32 --e2e-steps 2 --no-cpu-baseline > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 40 -c 1 -o gpurun_out/prof_sweep_$TAG -f python bench.py --steps 1 --warmup 1 --samples 16 --graphs 0 --e2e-steps 2 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_sweep_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv
ncu -i gpurun_out/prof_sweep_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_$TAG.csv 2>&1
python profiles/configs_bench.py --json gpurun_out/configs_$TAG.json > gpurun_out/configs_$TAG.txt 2>&1; cat gpurun_out/configs_$TAG.txt
