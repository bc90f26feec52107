( timeout 900 python -m pytest tests/test_gpu_benched_path.py -x -q -m gpu ) 2>&1 | tail -3
run() { for c in "C4" "C5" "C3" "L=4096 x 1" "x 4 replicas" "C2"; do timeout 300 python profiles/configs_bench.py --samples 128 --only "$c" 2>&1 | tail -1; done; }
echo "=== HEAD defaults (PDL, sweeps above pyramids with node priorities, up to 64 slots, 64-sample graphs)"; run
echo "=== round-2 settings before this session (MCRG_PDL=0 MCRG_SLOTS=4 MCRG_GRAPH_CHUNK=16 MCRG_GRAPH_PRIO=0 MCRG_PYR_PRIO=hi)"; MCRG_PDL=0 MCRG_SLOTS=4 MCRG_GRAPH_CHUNK=16 MCRG_GRAPH_PRIO=0 MCRG_PYR_PRIO=hi run
echo "=== HEAD defaults again"; run
python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; python -c "
import json; d=json.load(open('gpurun_out/bench_head.json')); print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'other',d['other_schedules_per_gpu'],{k:v['value'] for k,v in d['other_configs'].items()})"
