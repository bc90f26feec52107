run() { for c in "C4" "C5" "C3"; do timeout 300 python profiles/configs_bench.py --samples 256 --only "$c" 2>&1 | tail -1; done; }
echo "=== HEAD (PDL=3)"; run
echo "=== MCRG_SPLIT_MEASURE=1 (measure0 + pyramid low priority)"; MCRG_SPLIT_MEASURE=1 run
echo "=== MCRG_SPLIT_MEASURE=1 MCRG_PYR_PRIO=same"; MCRG_SPLIT_MEASURE=1 MCRG_PYR_PRIO=same run
echo "=== MCRG_SPLIT_MEASURE=1 MCRG_PYR_PRIO=hi"; MCRG_SPLIT_MEASURE=1 MCRG_PYR_PRIO=hi run
( MCRG_SPLIT_MEASURE=1 timeout 600 python -m pytest tests/test_gpu_benched_path.py -x -q -m gpu -k "oracle" ) 2>&1 | tail -2
