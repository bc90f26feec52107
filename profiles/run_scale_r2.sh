# Run with gpurun --gpus 8: two-GPU exactness tests, then bench.py under torchrun at 8 / 4 / 2 GPUs of one box for C4 (headline;
# e2e + host roofline per N) and C5 (one 16384^2 replica per GPU, 8 levels); outputs in gpurun_out/.
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1; lscpu > gpurun_out/lscpu_8gpu.txt; nproc
timeout 600 python -m pytest tests/test_gpu_multidevice.py -q -v > gpurun_out/pytest_multidevice_r2.log 2>&1; echo "multidevice rc=$?"; tail -6 gpurun_out/pytest_multidevice_r2.log
for cfg in C4 C5; do
  for n in 8 4 2; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --config $cfg --steps 10 --warmup 3 > gpurun_out/scale_r2_${cfg}_n${n}.json 2> gpurun_out/scale_r2_${cfg}_n${n}.err
    echo "$cfg n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_r2_${cfg}_n${n}.json').read().strip().splitlines()[-1])
    r=d['e2e']['host_roofline']
    print(d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['e2e']['path'], {k:round(v,1) for k,v in d['e2e']['variants'].items()}, 'roofline', round(r['value'],1), 'read/pack/dma', round(r['read_GBps_per_rank'],1), round(r['pack_GBps_per_rank'],1), round(r['dma_GBps_per_rank'],1), d['e2e']['host_threads'], d['e2e']['host_threads_how'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/scale_r2_${cfg}_n${n}.err').read()[-1500:])
PY
  done
done
