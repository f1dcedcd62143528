python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_7.log
tail -4 gpurun_out/r02_pytest_gpu_7.log
for c in c4 c5 c3; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02b_bench_$c.json 2>> gpurun_out/r02b.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r02b_bench_$c.json')); print('$c', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s rollout', round(d['roofline']['kernel_ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"; done
