# Run with gpurun --gpus 8: bench.py under torchrun at 8 and 2 GPUs of one box for C4 (HEAD at the end of round 2: PDL, priorities, 64 slots)
for n in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config C4 --steps 10 --warmup 3 > gpurun_out/scale_r2b_C4_n${n}.json 2> gpurun_out/scale_r2b_C4_n${n}.err
  echo "C4 n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_r2b_C4_n${n}.json').read().strip().splitlines()[-1])
    r=d['e2e']['host_roofline']
    print(d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['e2e']['path'], {k:round(v,1) for k,v in d['e2e']['variants'].items()}, 'roofline', round(r['value'],1), d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/scale_r2b_C4_n${n}.err').read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r2b_C5_n8.json 2> gpurun_out/scale_r2b_C5_n8.err; python -c "
import json; d=json.loads(open('gpurun_out/scale_r2b_C5_n8.json').read().strip().splitlines()[-1]); print('C5 n=8 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
