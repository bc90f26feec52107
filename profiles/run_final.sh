# final round-2 pass on one B200: sanitizer on HEAD's kernels, then the full profile script
( timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize.py all ) > gpurun_out/sanitizer_memcheck_r2.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_r2.txt
( timeout 1200 compute-sanitizer --tool racecheck python profiles/sanitize.py sweep ) > gpurun_out/sanitizer_racecheck_r2.txt 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_r2.txt
TAG=r2_final bash profiles/run_profiles.sh
