# usage (under gpurun): bash tools/dev/exp1.sh  -- parity (quick_check + pytest -m gpu) of the default build, then A/B of build_ab/*.so
mkdir -p gpurun_out
timeout 200 python tools/quick_check.py > gpurun_out/quick.log 2>&1; echo "quick rc=$?"; tail -2 gpurun_out/quick.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
bash tools/dev/ab.sh "${WLS:-c5 c1 c4}" $LIBS
for lib in $LIBS; do
  ZMESH_B200_LIB=$PWD/build_ab/$lib timeout 200 python bench.py --workload c5z --no-cpu --no-e2e --steps 8 > gpurun_out/ab_${lib}_c5z.json 2> gpurun_out/ab_${lib}_c5z.err
  python -c "
import json
d=json.load(open('gpurun_out/ab_${lib}_c5z.json'))
print('$lib', 'c5z', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
"
done
