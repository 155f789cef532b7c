# compute-sanitizer runs of the small sanity script (kept under profiles/)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_check.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/quick_check.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
