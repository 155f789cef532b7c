# round 2 final measurement set (1 GPU): tests, the default bench as the driver runs it, reference arm, c1 / c4 lines,
# ncu launch list of the default bench, ncu --set full of the final kernels (c5s, c1), DRAM traffic of the c5 kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 rc=$?"; tail -c 300 gpurun_out/r02_bench_c5.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c5_reference_arm.json 2> gpurun_out/r02_bench_c5_reference_arm.err; echo "ref rc=$?"
for w in c1 c4; do timeout 300 python bench.py --workload $w --steps 10 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c5.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_c5.log 2>&1; echo "launch list rc=$?"
bash tools/dev/prof.sh r02final c5s c1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_classify|k_emit' -s 9 -c 3 \
  --csv --log-file gpurun_out/r02_traffic_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/traffic_c5.log 2>&1; echo "traffic rc=$?"
