set -x
python tests/experiments/grasp_parity.py > gpurun_out/r02_grasp_parity.log 2>&1
python tools/quick_time.py > gpurun_out/r02_quick_time_a.log 2>&1
M3P2I_TEAM_ALIGN=0 python tools/quick_time.py > gpurun_out/r02_quick_time_noalign.log 2>&1
M3P2I_LANES=16 python tools/quick_time.py > gpurun_out/r02_quick_time_l16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_team -s 3 -c 1 -o gpurun_out/r02_grasp python tools/grasp_case.py > gpurun_out/ncu_grasp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_team -s 3 -c 1 -o gpurun_out/r02_rest python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rest.log 2>&1
cat gpurun_out/r02_grasp_parity.log gpurun_out/r02_quick_time_a.log gpurun_out/r02_quick_time_noalign.log gpurun_out/r02_quick_time_l16.log
