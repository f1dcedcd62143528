"""Drives the UNMODIFIED reference planner code (m3p2i_aip.planners.motion_planner.{m3p2i,mppi,cost_functions},
utils.{skill_utils,mppi_utils}) exactly as scripts/reactive_tamp.py does (run_tamp / dynamics / running_cost,
reactive_tamp.py:43-73) over this repo's sim facade backed by the CPU oracle integrator (PhysX cannot be installed).
TEST INFRASTRUCTURE ONLY: used by tests/golden/make_golden.py (fixtures), tests/test_reference_direct.py and bench.py's
CPU legs (cpu_baseline, --impl reference).

Where the reference comes from: `baseline/_ref` (pip install --no-deps --target baseline/_ref of the reference checkout:
git-ignored, travels to the GPU box) or, in the build container, /root/reference/src. Nothing is modified or copied:
two sys.modules stubs stand in for packages that are not installable (`ghalton`, `isaacgym`), and the module-level name
`torch` seen by skill_utils is wrapped so that its hard-coded torch.zeros(..., device='cuda:0') (skill_utils.py:69)
allocates on the CPU.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "m3p2i-aip_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py as O  # noqa: E402
from m3p2i_b200 import scene as S  # noqa: E402
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as our_wrapper  # noqa: E402  (this repo's facade)

CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"]


def reference_dir():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "m3p2i_aip", "planners", "motion_planner", "mppi.py")):
            return d
    return None


def import_reference(src=None):
    """-> (reference m3p2i module, reference cost_functions module). The facade of this repo is imported first (above)
    and the reference package is then loaded under its own name from a separate module table."""
    src = src or reference_dir()
    if src is None:
        raise ImportError("the reference package is neither in baseline/_ref nor in /root/reference/src")
    saved = {k: v for k, v in sys.modules.items() if k == "m3p2i_aip" or k.startswith("m3p2i_aip.")}
    for k in saved:
        del sys.modules[k]
    gh = types.ModuleType("ghalton")
    gh.EA_PERMS = []
    gh.GeneralizedHalton = object
    ig = types.ModuleType("isaacgym")
    gymapi = types.ModuleType("isaacgym.gymapi")
    gymapi.SimParams = type("SimParams", (), {})
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    ig.gymapi, ig.gymtorch = gymapi, gymtorch
    stubs = {"ghalton": gh, "isaacgym": ig, "isaacgym.gymapi": gymapi, "isaacgym.gymtorch": gymtorch}
    had = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    sys.path.insert(0, src)
    try:
        from m3p2i_aip.planners.motion_planner import m3p2i as ref_m3p2i
        from m3p2i_aip.planners.motion_planner import cost_functions as ref_cost
        from m3p2i_aip.utils import skill_utils as ref_skill
    finally:
        sys.path.remove(src)
    assert os.path.abspath(ref_m3p2i.__file__).startswith(os.path.abspath(src)), ref_m3p2i.__file__

    class _TorchCPU:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def zeros(*a, **kw):
            kw.pop("device", None)
            return torch.zeros(*a, **kw)
    ref_skill.torch = _TorchCPU()
    ref = {k: v for k, v in sys.modules.items() if k == "m3p2i_aip" or k.startswith("m3p2i_aip.")}
    for k in ref:
        del sys.modules[k]
    sys.modules.update(saved)
    for k, v in had.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return ref_m3p2i, ref_cost


class Tamp:
    """scripts/reactive_tamp.py:21-73 without zerorpc/hydra and with a fixed task (no task planner)."""

    def __init__(self, cfg, ref_m3p2i, ref_cost):
        self.sim = our_wrapper.IsaacGymWrapper(cfg.isaacgym, cfg.env_type, num_envs=cfg.mppi.num_samples, viewer=False,
                                               device="cpu", cube_on_shelf=cfg.cube_on_shelf,
                                               backend_factory=O.Oracle.for_sim)
        self.cfg = cfg
        self.objective = ref_cost.Objective(cfg)
        self.motion_planner = ref_m3p2i.M3P2I(cfg, dynamics=self.dynamics, running_cost=self.running_cost)

    def dynamics(self, _, u, t=None):
        self.sim.set_dof_velocity_target_tensor(u)
        self.sim.step()
        states = torch.stack([self.sim.robot_pos[:, 0], self.sim.robot_vel[:, 0], self.sim.robot_pos[:, 1],
                              self.sim.robot_vel[:, 1]], dim=1)
        return states, u

    def running_cost(self, _):
        return self.objective.compute_cost(self.sim)

    def run_tamp(self, dof_state, root_state, task, goal, extra_step):
        self.sim._dof_state[:] = dof_state
        self.sim._root_state[:] = root_state
        self.sim.set_dof_state_tensor(self.sim._dof_state)
        self.sim.set_actor_root_state_tensor(self.sim._root_state)
        if extra_step:
            self.sim.step()  # PLANNER_AIF_PANDA.update_plan, task_planner.py:79
        self.motion_planner.update_gripper_command(task)
        self.objective.update_objective(task, goal)
        return self.motion_planner.command(self.sim._dof_state[0])


def time_reference(env, task, goal, K, T, dof, root, multi_modal=False, cube_on_shelf=False, steps=3, warmup=1,
                   threads=None, seed=0):
    """Wall time of the reference's own M3P2I.command() (+ Objective.compute_cost, the T-step Python loop, the softmin
    update, top-k, Savitzky-Golay) on the CPU, PhysX replaced by the oracle integrator behind the facade. The noise
    table is injected (planner.delta) so that the K * nu scipy spline loop of the first call is not timed.
    -> (list of per-command seconds, source directory of the reference)."""
    import time
    threads = threads or os.cpu_count() or 1
    O.set_threads(threads)
    torch.set_num_threads(threads)
    ref_m3p2i, ref_cost = import_reference()
    cfg = S.make_cfg(env, task, goal, K, T, multi_modal=multi_modal, cube_on_shelf=cube_on_shelf, device="cpu")
    torch.manual_seed(seed)
    tamp = Tamp(cfg, ref_m3p2i, ref_cost)
    nu = tamp.motion_planner.nu
    delta = np.random.default_rng(seed).standard_normal((K, T, nu)).astype(np.float32)
    tamp.motion_planner.delta = torch.from_numpy(delta)
    d, r = torch.from_numpy(np.asarray(dof, np.float32)), torch.from_numpy(np.asarray(root, np.float32))
    g = torch.tensor(np.asarray(goal, np.float32))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tamp.run_tamp(d, r, task, g, False)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tamp.sim.stop_sim()
    return times, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(ref_m3p2i.__file__))))
