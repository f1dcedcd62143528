"""Objective: the task-switched per-step cost plugin (reference: planners/motion_planner/cost_functions.py:5-36).

Same constructor, update_objective(task, goal) and compute_cost(sim) -> cost[K]. The arithmetic of every task
cost (navigation / push / pull / push_pull / reach / pick / place + collision, cost_functions.py:38-169) runs in
the CUDA kernels; inside the fused MPPI.command() it is evaluated in the rollout kernel and this object only
carries task and goal. compute_cost(sim) evaluates the same device function on the sim's K persistent envs.
"""
import numpy as np
import torch

from m3p2i_b200 import _abi as A


class Objective(object):
    def __init__(self, cfg):
        self.cfg = cfg
        self.multi_modal = cfg.multi_modal
        self.num_samples = cfg.mppi.num_samples
        self.half_samples = int(cfg.mppi.num_samples / 2)
        self.device = "cpu"
        self.pre_height_diff = getattr(cfg, "pre_height_diff", 0.0)
        self.tilt_cos_theta = 0.5
        self.task = None
        self.goal = None

    def update_objective(self, task, goal):
        if task not in A.TASK_IDS:
            raise ValueError(f"unknown task {task!r}; expected one of {sorted(A.TASK_IDS)}")
        self.task = task
        self.goal = goal if torch.is_tensor(goal) else torch.tensor(goal, dtype=torch.float32)

    def goal_array(self):
        return np.asarray(self.goal.detach().cpu().numpy() if torch.is_tensor(self.goal) else self.goal,
                          np.float32).ravel()

    def compute_cost(self, sim):
        if self.task is None:
            raise RuntimeError("update_objective(task, goal) must be called before compute_cost")
        if not getattr(sim, "_has_planner_cfg", False):
            # the sim was created from the isaacgym section alone (reactive_tamp.py:23-30); the cost kernels also
            # need kp_suction / multi_modal / pre_height_diff of the full configuration
            sim.attach_planner(self.cfg)
        sim._push()
        sim.backend.set_objective(self.task, self.goal_array(), None)
        cost = torch.from_numpy(sim.backend.sim_cost())
        sim._host_dirty = True  # the pull cost arms suction forces inside the envs
        return cost
