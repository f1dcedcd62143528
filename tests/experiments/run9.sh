# final round-2 evidence on one B200 (gpurun -- bash tests/experiments/run9.sh): tests, one bench line per config, the
# reference arm, the ncu launch list, ncu --set full of k_rollout_far (C4, C5) and k_rollout_team (grasp state), K sweep
set -x
O=gpurun_out
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > $O/r02_pytest_gpu.log; tail -3 $O/r02_pytest_gpu.log
python bench.py --steps 100 --warmup 10 > $O/r02_bench_n1_c4.json 2> $O/r02_bench.err; cut -c1-300 $O/r02_bench_n1_c4.json
for c in c1 c2 c3 c4_grasp c5; do python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline > $O/r02_bench_n1_$c.json 2>> $O/r02_bench.err; python -c "
import json,sys; d=json.load(open('$O/r02_bench_n1_$c.json')); print('$c', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s rollout', round(d['roofline']['kernel_ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), d.get('far_field',{}).get('near_samples'))"; done
python bench.py --impl reference --steps 5 --warmup 3 > $O/r02_bench_reference_arm_c4.json 2>> $O/r02_bench.err; cut -c1-300 $O/r02_bench_reference_arm_c4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_far -s 4 -c 1 -f -o $O/r02_far_c4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_far -s 4 -c 1 -f -o $O/r02_far_c5 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_team -s 3 -c 1 -f -o $O/r02_grasp python tools/grasp_case.py > /dev/null 2>&1
python tools/ksweep.py 1024 2048 4096 6144 8192 12288 16384 32768 65536 262144 > $O/r02_ksweep_auto.csv 2>&1; cat $O/r02_ksweep_auto.csv
tail -3 $O/r02_bench.err
