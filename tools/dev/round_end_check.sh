# usage (under gpurun): bash tools/dev/round_end_check.sh   -- GPU test suite with durations, then the ncu launch list of the default bench (c5)
mkdir -p gpurun_out
timeout 450 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_c5.log 2>&1; echo "ncu rc=$?"; tail -c 700 gpurun_out/launches_c5.log
