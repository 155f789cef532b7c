mkdir -p gpurun_out
timeout 100 python tools/quick_check.py > gpurun_out/quick.log 2>&1; echo "quick rc=$?"; tail -1 gpurun_out/quick.log
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in ${WLS:-c5 c1 c4}; do timeout 300 python bench.py --workload $w --no-cpu --no-e2e --steps 8 > gpurun_out/b_$w.json 2> gpurun_out/b_$w.err; python -c "
import json
d=json.load(open('gpurun_out/b_$w.json'))
print('$w', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
"; done
