# usage (under gpurun): bash tools/dev/final_run.sh <tag>   -- the measurement set kept under profiles/
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; tail -c 400 gpurun_out/bench_${TAG}_c5.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_c5_ref.json 2> gpurun_out/bench_${TAG}_c5_ref.err; tail -c 700 gpurun_out/bench_${TAG}_c5_ref.json
timeout 600 python bench.py --workload c1 > gpurun_out/bench_${TAG}_c1.json 2> gpurun_out/bench_${TAG}_c1.err; tail -c 300 gpurun_out/bench_${TAG}_c1.json
timeout 600 python bench.py --workload c4 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err; tail -c 300 gpurun_out/bench_${TAG}_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}_c1.csv \
  python bench.py --workload c1 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_${TAG}_c1.log 2>&1
timeout 800 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_classify|k_emit' -s 9 -c 3 \
  --csv --log-file gpurun_out/traffic_${TAG}_c5.csv python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/traffic_${TAG}_c5.log 2>&1
bash tools/dev/prof.sh $TAG c1 c5s
