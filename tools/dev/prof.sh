# usage (under gpurun): bash tools/dev/prof.sh <tag> [workloads...]
# ncu --set full with source counters for the three pipeline kernels of one timed step per workload
TAG=${1:-x}; shift
mkdir -p gpurun_out
for w in ${@:-c1 c5s}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_emit' -s 9 -c 3 \
    -f -o gpurun_out/prof_${TAG}_$w python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_${TAG}_$w.log 2>&1
  tail -2 gpurun_out/prof_${TAG}_$w.log
done
