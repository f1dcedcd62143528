"""Generates tests/golden/*.npz by running the UNMODIFIED reference planner code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

What is imported from the reference, unmodified, from /root/reference/src:
    m3p2i_aip.planners.motion_planner.m3p2i.M3P2I   (and mppi.MPPI)      -- sampling, rollout loop, softmin update
    m3p2i_aip.planners.motion_planner.cost_functions.Objective           -- per-step task costs
    m3p2i_aip.utils.skill_utils / mppi_utils                             -- suction, quaternion costs, cost_to_go
with two sys.modules stubs for packages that are not installable here (`ghalton`, `isaacgym`) and one shim:
skill_utils.calculate_suction hard-codes device='cuda:0' (skill_utils.py:69); the module-level name `torch` seen
by skill_utils is wrapped so that torch.zeros(..., device='cuda:0') allocates on the CPU. No reference source is
changed or copied.

The reference code is driven exactly as scripts/reactive_tamp.py drives it (run_tamp / dynamics / running_cost,
reactive_tamp.py:43-73) against the sim facade of this repo backed by the CPU oracle integrator (PhysX is not
available). Each case records the inputs of every tick (real state, noise table, task, goal) and the reference's
outputs (action, means, best trajectories, weights, cost_total, top-k, states, actions, beta). tests/ then
require the C oracle (and through it the CUDA path) to reproduce these outputs from the same inputs.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200"))

import oracle_py as O  # noqa: E402
from m3p2i_b200 import scene as S  # noqa: E402
from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as our_wrapper  # noqa: E402  (this repo's facade)

from reference_rig import Tamp, import_reference as _import_reference  # noqa: E402  (oracle/reference_rig.py)

REF_SRC = "/root/reference/src"


def import_reference():
    """The goldens are generated from the reference checkout itself."""
    return _import_reference(REF_SRC)


CASES = [
    dict(name="nav_k200_t12", env="point_env", task="navigation", goal=[-3.0, 3.0], K=200, T=12, calls=4),
    dict(name="nav_obstacle", env="point_env", task="navigation", goal=[-3.0, 3.0], K=128, T=12, calls=3,
         robot=[-1.55, 1.55]),
    dict(name="push_k256_t20", env="point_env", task="push", goal=[-1.0, -1.0], K=256, T=20, calls=4,
         robot=[0.2, 2.45]),
    dict(name="pull_k256_t20", env="point_env", task="pull", goal=[0.0, 0.0], K=256, T=20, calls=4,
         robot=[0.0, 1.55]),
    dict(name="push_pull_mm", env="point_env", task="push_pull", goal=[-3.75, -3.75], K=256, T=20, calls=4,
         multi_modal=True, robot=[0.3, 2.4]),
    dict(name="panda_reach", env="panda_env", task="reach", K=64, T=12, calls=3),
    dict(name="panda_reach_mm_shelf", env="panda_env", task="reach", K=64, T=12, calls=3, multi_modal=True,
         cube_on_shelf=True),
    dict(name="panda_pick", env="panda_env", task="pick", K=64, T=12, calls=3, extra_step=True, grasp=True),
    dict(name="panda_place", env="panda_env", task="place", K=32, T=12, calls=2),
]


def grasp_pose(scene):
    """Joint configuration whose fingers straddle cubeA on the table (found by damped least squares on the oracle FK)."""
    q = np.array([0, 0.3, 0, -2.2, 0, 2.5, 0.785, 0.04, 0.04], np.float64)
    target = np.array([0.2, -0.2, 1.06 + 0.045])
    for _ in range(200):
        def f(qq):
            L = O.panda_fk(scene, qq.astype(np.float32))
            ee = 0.5 * (L[0, :3] + L[1, :3]).astype(np.float64)
            R = np.zeros((3, 3))
            x, y, z, w = L[0, 3:7].astype(np.float64)
            zax = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])
            yax = np.array([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)])
            return np.concatenate([ee - target, 0.2 * (zax - np.array([0, 0, -1.0])), 0.2 * (yax[2:3])])
        e = f(q)
        if np.linalg.norm(e) < 1e-5:
            break
        J = np.zeros((e.size, 7))
        for j in range(7):
            dq = q.copy()
            dq[j] += 1e-4
            J[:, j] = (f(dq) - e) / 1e-4
        q[:7] -= J.T @ np.linalg.solve(J @ J.T + 1e-6 * np.eye(e.size), e)
    return q.astype(np.float32)


def run_case(case, ref_m3p2i, ref_cost, seed):
    env = case["env"]
    K, T = case["K"], case["T"]
    mm = case.get("multi_modal", False)
    shelf = case.get("cube_on_shelf", False)
    cfg = S.make_cfg(env, case["task"], case.get("goal"), K, T, multi_modal=mm, cube_on_shelf=shelf, device="cpu")
    torch.manual_seed(seed)  # MPPI.__init__ draws self.U (unused in halton-spline mode)
    tamp = Tamp(cfg, ref_m3p2i, ref_cost)
    nu = tamp.motion_planner.nu
    rng = np.random.default_rng(seed)
    delta = rng.standard_normal((K, T, nu)).astype(np.float32)
    tamp.motion_planner.delta = torch.from_numpy(delta.copy())

    actors = S.default_actors(env)
    dof = S.initial_dof_state(actors).copy()
    root = S.initial_root_state(actors, shelf).copy()
    if "robot" in case:
        dof[0], dof[2] = case["robot"]
    if case.get("grasp"):
        q = grasp_pose(tamp.sim.scene)
        dof[0::2] = q
        dof[14], dof[16] = 0.027, 0.027  # fingers nearly closed on the 5 cm cube
    goal = case.get("goal")
    if env == "panda_env":
        if case["task"] == "pick":
            cube_b = root[S.actor_index(actors, "cubeB")]
            goal = np.concatenate([cube_b[:3] + np.array([0, 0, 0.055], np.float32), cube_b[3:7]]).tolist()
        else:
            goal = [0.0] * 7
    # the "real world": one env of the same integrator, advanced with the first action of every command
    real = O.Oracle(S.build_config(S.sim_only_cfg(env, 1, cfg.isaacgym)), tamp.sim.scene)

    out = {"delta": delta, "goal": np.asarray(goal, np.float32), "task": case["task"], "env": env, "K": K, "T": T,
           "multi_modal": mm, "cube_on_shelf": shelf, "extra_step": bool(case.get("extra_step", False)),
           "calls": case["calls"], "nu": nu}
    for i in range(case["calls"]):
        out[f"dof_{i}"] = dof.copy()
        out[f"root_{i}"] = root.copy()
        action = tamp.run_tamp(torch.from_numpy(dof), torch.from_numpy(root), case["task"],
                               torch.tensor(goal, dtype=torch.float32), out["extra_step"])
        mp = tamp.motion_planner
        out[f"action_{i}"] = action.numpy().astype(np.float32)
        out[f"mean_action_{i}"] = mp.mean_action.numpy().copy()
        out[f"weights_{i}"] = mp.weights.numpy().copy()
        out[f"cost_total_{i}"] = mp.cost_total.numpy().copy()
        out[f"states_{i}"] = mp.states.numpy().copy()
        out[f"actions_{i}"] = mp.actions.numpy().copy()
        out[f"top_idx_{i}"] = mp.top_idx.numpy().astype(np.int32)
        out[f"top_values_{i}"] = mp.top_values.numpy().copy()
        out[f"top_trajs_{i}"] = mp.top_trajs.numpy().copy()
        out[f"beta_{i}"] = np.float64(mp.beta)
        if mm:
            out[f"mean_action_1_{i}"] = mp.mean_action_1.numpy().copy()
            out[f"mean_action_2_{i}"] = mp.mean_action_2.numpy().copy()
            out[f"best_traj_1_{i}"] = mp.best_traj_1.numpy().copy()
            out[f"best_traj_2_{i}"] = mp.best_traj_2.numpy().copy()
            out[f"weights_1_{i}"] = mp.weights_1.numpy().copy()
            out[f"weights_2_{i}"] = mp.weights_2.numpy().copy()
            out[f"pull_preference_{i}"] = np.int32(mp.get_pull_preference())
        else:
            out[f"best_traj_{i}"] = mp.best_traj.numpy().copy()
        # advance the real world a few ticks with the commanded first action
        real.set_state(dof, root)
        real.sim_set_velocity_target(out[f"action_{i}"][0:1])
        for _ in range(3):
            real.sim_step()
        d, r, _, _ = real.sim_read()
        dof, root = d[0].copy(), r[0].copy()
    return out


def main():
    ref_m3p2i, ref_cost = import_reference()
    O.set_threads(1)
    for n, case in enumerate(CASES):
        out = run_case(case, ref_m3p2i, ref_cost, seed=1000 + n)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        ct = out[f"cost_total_{case['calls'] - 1}"]
        print(f"{case['name']:24s} -> {os.path.getsize(path) / 1024:7.1f} KiB  cost_total[min,max]=({ct.min():.3f}, "
              f"{ct.max():.3f})  n_coll={(out['states_0'].shape[0])}")


if __name__ == "__main__":
    main()
