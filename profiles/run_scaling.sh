# Run with gpurun --gpus 8: bench.py under torchrun at 8, 4, 2 GPUs of one box, then 1 GPU; outputs in gpurun_out/.
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/bench_n${n}_final.json 2> gpurun_out/bench_n${n}_final.err
  echo "rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_n${n}_final.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_final.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_final.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
