mkdir -p gpurun_out; timeout 300 python tools/quick_check.py > gpurun_out/quick.log 2>&1; tail -3 gpurun_out/quick.log; if grep -q "random_u32_dense_tma ok" gpurun_out/quick.log; then timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log; for w in ${WL:-c1 c4 c5}; do timeout 400 python bench.py --workload $w --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json'))
print('$w', round(d['value']), 'MVx/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, 'pipe frac', round(d['roofline']['pipeline']['frac'],4))
"; done; else compute-sanitizer --tool memcheck python tools/quick_check.py 2>&1 | grep -v "Host Frame" | head -40; fi
