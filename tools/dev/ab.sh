# usage (under gpurun): bash tools/dev/ab.sh "<workloads>" lib1.so lib2.so ...   -- A/B of development builds
WLS=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  for w in $WLS; do
    ZMESH_B200_LIB=$PWD/build_ab/$lib timeout 400 python bench.py --workload $w --no-cpu --no-e2e --steps 8 > gpurun_out/ab_${lib}_$w.json 2> gpurun_out/ab_${lib}_$w.err
    python -c "
import json
d=json.load(open('gpurun_out/ab_${lib}_$w.json'))
print('$lib', '$w', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
"
  done
done
