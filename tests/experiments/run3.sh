python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu_3.log
tail -4 gpurun_out/r02_pytest_gpu_3.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
cat gpurun_out/r02_bench_c4.json
for c in c1 c2 c3 c4_grasp c5; do python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2>> gpurun_out/r02_bench_c4.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench_$c.json')); print('$c', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s rollout', round(d['roofline']['kernel_ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"; done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_c4.json 2>> gpurun_out/r02_bench_c4.err; cat gpurun_out/r02_bench_reference_c4.json
tail -5 gpurun_out/r02_bench_c4.err
