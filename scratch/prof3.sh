ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 40 -c 1 -o gpurun_out/prof_sweep_r1_final -f python bench.py --steps 1 --warmup 1 --samples 16 --graphs 0 --no-cpu-baseline > gpurun_out/ncu_r1_final.log 2>&1
ncu -i gpurun_out/prof_sweep_r1_final.ncu-rep --page raw --csv > gpurun_out/raw_r1_final.csv
ncu -i gpurun_out/prof_sweep_r1_final.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_r1_final.csv 2>&1
head -3 gpurun_out/raw_r1_final.csv | cut -c1-200
