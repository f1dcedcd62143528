"""Closed-loop reach driven by the ORACLE planner; every tick the CUDA planner gets the oracle's planner state and the
same real state and its command is compared (action, per-sample costs).   python tests/experiments/reach_sync.py [lanes] [ticks]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT]
import oracle_py as O
from helpers import make_backend
from m3p2i_b200 import _abi as A, native, scene as S
from m3p2i_aip.utils import mppi_utils
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as wrapper
O.set_threads(os.cpu_count())
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 120
sync = int(os.environ.get("SYNC", "1"))
K, H = 1024, 16
cfg = S.make_cfg("panda_env", "reach", None, K, H)
cfg.mppi.lanes_per_sample = lanes
table = mppi_utils.halton_spline_table(K, H, 9)
o = make_backend(O.Oracle, cfg); n = make_backend(native.NativePlanner, cfg)
for b in (o, n):
    b.set_noise_table(table)
real = wrapper.IsaacGymWrapper(cfg.isaacgym, "panda_env", num_envs=1, device="cpu", backend_factory=O.Oracle.for_sim)
for _ in range(30):
    real.step()
for i in range(ticks):
    dof = real._dof_state[0].clone().numpy().astype(np.float32); root = real._root_state.clone().numpy().astype(np.float32)
    for b in (o, n):
        b.set_state(dof, root); b.set_objective("reach", np.zeros(7, np.float32), "open")
    if sync: n.set_planner_state(o.get_planner_state())
    a_n, c_n, _ = n.command(); a_o, c_o, _ = o.command()
    a_n, a_o = np.array(a_n), np.array(a_o)
    ch_n, ch_o = n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON)
    bad = ~np.isclose(ch_n, ch_o, rtol=1e-3, atol=1e-3)
    use = a_n if os.environ.get("DRIVE") == "native" else a_o
    real.set_dof_velocity_target_tensor(torch.tensor(use[0]).view(1, -1)); real.step()
    ee = 0.5 * (real.get_actor_link_by_name("panda", "panda_leftfinger")[0, :3] + real.get_actor_link_by_name("panda", "panda_rightfinger")[0, :3])
    cube = real.get_actor_link_by_name("cubeA", "box")[0, :3]
    d = float(torch.linalg.norm(ee - cube))
    if i % 4 == 0 or bad.any(1).mean() > 0.02:
        print(f"tick {i}: ee-cube {d:.4f} samples differing {bad.any(1).mean():.4f} max|dc| {np.abs(ch_n - ch_o).max():.3e} action diff {np.abs(a_n - a_o).max():.2e} "
              f"(first {np.abs(a_n[0] - a_o[0]).max():.2e}) cost_total diff {np.abs(np.array(c_n) - np.array(c_o)).max():.2e}", flush=True)
