# ncu --set full of one k_sweep0 launch for each library under gpurun_ab/ named in $LIBS; prints the metrics that matter here
for name in $LIBS; do
  MCRG_LIB=$PWD/gpurun_ab/lib$name.so ncu --set full --clock-control none --import-source on -k regex:k_sweep0 -s 34 -c 1 -o gpurun_out/ncu_ab_$name -f python profiles/sweep_driver.py ${MODE:-sweep} > gpurun_out/ncu_ab_$name.log 2>&1
  ncu -i gpurun_out/ncu_ab_$name.ncu-rep --page raw --csv > gpurun_out/ncu_ab_$name.csv 2>/dev/null
  python - "$name" <<'PY'
import csv,sys
name=sys.argv[1]
rows=list(csv.reader(open(f'gpurun_out/ncu_ab_{name}.csv')))
hdr,vals=rows[0],rows[-1]
d=dict(zip(hdr,vals))
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','smsp__warps_eligible.avg.per_cycle_active']
print('==',name)
for k in keys:
    if k in d: print(f'  {k:75s} {d[k]}')
for k in sorted(d):
    if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k:
        try:
            if float(d[k])>0.15: print(f'  {k:75s} {d[k]}')
        except: pass
for k in sorted(d):
    if 'bank_conflict' in k or 'register' in k and 'conflict' in k: print(f'  {k:75s} {d[k]}')
PY
done
