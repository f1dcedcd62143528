"""Reach hover state (fingers open astride cubeA, a few cm above the pre-grasp pose): CUDA vs oracle on the states a
closed-loop oracle episode passes through.   python tests/experiments/hover_parity.py [lanes]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT]
import oracle_py as O
from helpers import make_backend
from m3p2i_b200 import _abi as A, native, scene as S
from m3p2i_aip.planners.motion_planner import m3p2i
from m3p2i_aip.planners.motion_planner.cost_functions import Objective
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as wrapper
O.set_threads(os.cpu_count())
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K, H = 1024, 16
factory = O.Oracle.for_sim


class Tamp:
    def __init__(self, cfg):
        self.sim = wrapper.IsaacGymWrapper(cfg.isaacgym, cfg.env_type, num_envs=cfg.mppi.num_samples, device="cpu", backend_factory=factory)
        self.objective = Objective(cfg)
        self.mp = m3p2i.M3P2I(cfg, dynamics=self.dynamics, running_cost=self.running_cost)
    def dynamics(self, _, u, t=None): raise AssertionError
    def running_cost(self, _): raise AssertionError
    def run_tamp(self, dof, root, task, goal):
        self.sim._dof_state[:] = dof; self.sim._root_state[:] = root
        self.sim.set_dof_state_tensor(self.sim._dof_state); self.sim.set_actor_root_state_tensor(self.sim._root_state)
        self.mp.update_gripper_command(task); self.objective.update_objective(task, goal)
        return self.mp.command(self.sim._dof_state[0])[0]


cfg = S.make_cfg("panda_env", "reach", None, K, H)
cfg.mppi.sampling_method = "halton"
tamp = Tamp(cfg)
real = wrapper.IsaacGymWrapper(cfg.isaacgym, "panda_env", num_envs=1, device="cpu", backend_factory=factory)
for _ in range(30):
    real.step()
states = []
for i in range(104):
    a = tamp.run_tamp(real._dof_state.clone(), real._root_state.clone(), "reach", torch.zeros(7))
    real.set_dof_velocity_target_tensor(a.view(1, -1)); real.step()
    if i >= 70 and i % 4 == 0:
        states.append((i, real._dof_state[0].clone().numpy().astype(np.float32), real._root_state.clone().numpy().astype(np.float32)))
cfg2 = S.make_cfg("panda_env", "reach", None, K, H)
cfg2.mppi.lanes_per_sample = lanes
for i, dof, root in states:
    o = make_backend(O.Oracle, cfg2, noise_mode=A.NOISE_PHILOX, seed=11)
    n = make_backend(native.NativePlanner, cfg2, noise_mode=A.NOISE_PHILOX, seed=11)
    for b in (o, n):
        b.set_state(dof, root); b.set_objective("reach", np.zeros(7, np.float32), "open")
    a_n, _, _ = n.command(); a_o, _, _ = o.command()
    ch_n, ch_o = n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON)
    d = np.abs(ch_n - ch_o)
    bad = ~np.isclose(ch_n, ch_o, rtol=1e-3, atol=1e-3)
    flips = ((ch_n > 900) != (ch_o > 900))
    print(f"tick {i}: samples differing {bad.any(1).mean():.4f} (collision flips {flips.any(1).mean():.4f}), median |dc| {np.median(d):.2e}, "
          f"99% {np.quantile(d, 0.99):.2e}, mean cost n {ch_n.mean():.3f} o {ch_o.mean():.3f}, coll steps n {int((ch_n > 900).sum())} o {int((ch_o > 900).sum())}, "
          f"action max diff {np.abs(a_n - a_o).max():.2e}", flush=True)
    o.close(); n.close()
