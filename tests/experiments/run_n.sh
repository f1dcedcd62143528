# multi-GPU evidence (gpurun --gpus N -- bash tests/experiments/run_n.sh N): C5 (multi-modal shelf reach) and C4 (pick),
# peer-memory exchange; C5 also with NCCL; the sharded-vs-unsharded check over the real multi-process mapping
N=$1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 100 --warmup 10 > $O/r02_bench_n${N}_c5.json 2> $O/r02_bench_n${N}.err
$TR bench.py --gpus $N --config c4 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_bench_n${N}_c4.json 2>> $O/r02_bench_n${N}.err
$TR bench.py --gpus $N --exchange nccl --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_bench_n${N}_c5_nccl.json 2>> $O/r02_bench_n${N}.err
EXCHANGE=peer $TR tools/nccl_check.py > $O/r02_peer_check_${N}gpu.log 2>&1
for f in c5 c4 c5_nccl; do python -c "
import json; d=json.load(open('$O/r02_bench_n${N}_$f.json')); print('$f N=$N', round(d['ms_per_step'],4), 'ms', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), 'eff', d.get('weak_scaling_efficiency_same_workload'), 'wait', d.get('peer_wait_ms',{}).get('wait_costs_mean_over_ranks'))"; done
tail -2 $O/r02_peer_check_${N}gpu.log; tail -3 $O/r02_bench_n${N}.err
