# A/B of programmatic dependent launch on one B200: MCRG_PDL=0 against the trigger positions (MCRG_PDL_TRIG = 0 top, 1 before the
# half-sweeps, 2 before the store, 3 at exit)
run() { for c in "L=4096 x 1" "C5" "C4" "C3"; do timeout 300 python profiles/configs_bench.py --only "$c" 2>&1 | tail -1; done; }
echo "=== MCRG_PDL=0"; MCRG_PDL=0 run
for t in 0 1 2 3; do echo "=== MCRG_PDL=1 MCRG_PDL_TRIG=$t"; MCRG_PDL_TRIG=$t run; done
echo "=== MCRG_PDL=0"; MCRG_PDL=0 run
