# new full-size config tests, then the kept bench lines of the non-default configs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "config3_full or config5_slab or config4_slab" --durations=5 > gpurun_out/pytest_cfg.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_cfg.log
for w in ${WLS:-c2a c2b c3 c4}; do timeout 600 python bench.py --workload $w --steps 8 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_$w.json'))
print('$w', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), 'parity', d['parity_on_sample'], d['details'])
"; tail -2 gpurun_out/r02_bench_$w.err; done
