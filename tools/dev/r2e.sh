# N-GPU bench (usage under gpurun --gpus N: N=2 bash tools/dev/r2e.sh)
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ${EXTRA} > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_c5_n$N.json; tail -5 gpurun_out/bench_c5_n$N.err
ZM_SHARD_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-parity > gpurun_out/bench_c5_n${N}_timing.json 2> gpurun_out/bench_c5_n${N}_timing.err; echo "rc=$?"; tail -3 gpurun_out/bench_c5_n${N}_timing.err
