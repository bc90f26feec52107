set -x
python profiles/configs_bench.py --json gpurun_out/configs_r1g.json > gpurun_out/configs_r1g.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 1 --samples 32 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 60 -c 1 -o gpurun_out/prof_sweep_r1g -f python bench.py --steps 1 --warmup 1 --samples 16 --graphs 0 --no-cpu-baseline > gpurun_out/ncu_r1g.log 2>&1
ncu -i gpurun_out/prof_sweep_r1g.ncu-rep --page raw --csv > gpurun_out/raw_r1g.csv
ncu -i gpurun_out/prof_sweep_r1g.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_r1g.csv 2>&1
cat gpurun_out/configs_r1g.txt
