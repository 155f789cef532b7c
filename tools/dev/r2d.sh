# bench.py end to end on the GPU box: default workload (c5) with cpu_baseline + e2e, the reference arm, c1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; tail -c 1500 gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c5_ref.json 2> gpurun_out/bench_c5_ref.err; echo "ref rc=$?"; tail -c 900 gpurun_out/bench_c5_ref.json; tail -3 gpurun_out/bench_c5_ref.err
timeout 600 python bench.py --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"; tail -c 1200 gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
