"""Host-side skill helpers used outside the rollout (reference: utils/skill_utils.py).

Inside MPPI.command() the suction model and the quaternion costs run in the CUDA kernels. What remains on the host
is what scripts/sim.py calls on the K=1 "real world" env every tick: the suction condition and force
(skill_utils.py:36-94) and the wall-clock pacing helper (:25-33).
"""
import time

import torch

from m3p2i_aip.utils.mppi_utils import bspline  # noqa: F401  (the reference defines it here, skill_utils.py:9-22)


def time_tracking(t, cfg):
    """Sleep up to the sim dt and report the real-time factor (skill_utils.py:25-33)."""
    actual_dt = time.time() - t
    rt = cfg.isaacgym.dt / actual_dt
    if rt > 1.0:
        time.sleep(cfg.isaacgym.dt - actual_dt)
        actual_dt = time.time() - t
        rt = cfg.isaacgym.dt / actual_dt
    print("FPS: {:.3f}".format(1 / actual_dt), "RT: {:.3f}".format(rt))
    return time.time()


def check_suction_condition(cfg, sim, action):
    """Suction is possible when the task pulls, the robot is within 0.6 m of the block and the commanded velocity
    points away from it (skill_utils.py:47-56)."""
    if cfg.task not in ['pull', 'push_pull'] or not cfg.suction_active:
        return False
    dir_robot_block = (sim.robot_pos - sim.get_actor_position_by_name("box")[:, :2]).squeeze(0)
    action_align_pull = torch.sum(torch.as_tensor(action, dtype=torch.float32).view(-1)[:2] * dir_robot_block).item()
    dis_robot_block = torch.linalg.norm(dir_robot_block)
    return bool(dis_robot_block < 0.6 and action_align_pull > 0)


def calculate_suction(cfg, sim):
    """forces [num_envs, bodies_per_env, 3]: -kp * unit(robot->block) on the block row, the opposite on the robot
    (last) row, when 1/dist exceeds 1.5 (one env) or 1.8 (rollout envs); clamped to +-500 (skill_utils.py:59-94)."""
    dir_vector = sim.get_actor_position_by_name("box")[:, :2] - sim.robot_pos
    magnitude = (1 / torch.linalg.norm(dir_vector, dim=1)).reshape([sim.num_envs, 1])
    unit_force = dir_vector * magnitude
    forces = torch.zeros((sim.num_envs, sim.bodies_per_env, 3), dtype=torch.float32)
    mask = (magnitude > (1.5 if sim.num_envs == 1 else 1.8)).reshape(sim.num_envs)
    block_index = int(sim._get_actor_index_by_name("box"))
    forces[mask, block_index, 0] = -cfg.kp_suction * unit_force[mask, 0]
    forces[mask, block_index, 1] = -cfg.kp_suction * unit_force[mask, 1]
    forces[mask, -1, 0] = cfg.kp_suction * unit_force[mask, 0]
    forces[mask, -1, 1] = cfg.kp_suction * unit_force[mask, 1]
    return torch.clamp(forces, min=-500, max=500)


def check_and_apply_suction(cfg, sim, action, verbose=False):
    """What scripts/sim.py:50 calls on the real env each tick (skill_utils.py:36-44)."""
    applied = False
    if check_suction_condition(cfg, sim, action):
        sim.apply_rigid_body_force_tensors(calculate_suction(cfg, sim))
        applied = True
    if verbose:
        print("suction!!!" if applied else "no suction...")
    return applied


# ---------------------------------------------------------------------------------------------------------------
# Quaternion orientation costs (skill_utils.py:140-180, 224-289). Inside the rollout they are evaluated by the CUDA
# kernels (panda_env.cuh); these host versions serve callers outside the hot path, e.g. the reference's task planner
# (task_planner.py:61), on small tensors.
def quaternion_rotation_matrix(Q):
    """[N,4] quaternions (x, y, z, w) -> [N,3,3] rotation matrices, same element formulas as the reference."""
    x, y, z, w = Q[:, 0], Q[:, 1], Q[:, 2], Q[:, 3]
    rows = (2 * (w * w + x * x) - 1, 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 2 * (w * w + y * y) - 1, 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 2 * (w * w + z * z) - 1)
    return torch.stack(rows, dim=1).reshape(Q.shape[0], 3, 3)


def _min_axis_cost(axis, R):
    """min over the three columns c of R of 1 - |<axis, c>|, per batch row."""
    return (1 - torch.abs(torch.einsum("ni,nij->nj", axis, R))).min(dim=1)[0]


def get_general_ori_cube2goal(cube_quaternion, goal_quatenion):
    """Alignment of the goal's x and y axes with any cube axis (invariant to flipped cubes)."""
    C, G = quaternion_rotation_matrix(cube_quaternion), quaternion_rotation_matrix(goal_quatenion)
    return _min_axis_cost(G[:, :, 0], C) + _min_axis_cost(G[:, :, 1], C)


def get_general_ori_ee2cube(ee_quaternion, cube_quaternion, tilt_value=0):
    """End-effector z axis perpendicular to a cube face (or at `tilt_value` to the cube axis closest to world x, taken
    from the FIRST row, skill_utils.py:275-279) plus y-axis alignment."""
    E, C = quaternion_rotation_matrix(ee_quaternion), quaternion_rotation_matrix(cube_quaternion)
    if tilt_value == 0:
        cost_z = _min_axis_cost(E[:, :, 2], C)
    else:
        sel = int(torch.argmax(torch.abs(C[0, 0, :])))
        cost_z = torch.abs(tilt_value - torch.sum(E[:, :, 2] * C[:, :, sel], dim=1))
    return cost_z + _min_axis_cost(E[:, :, 1], C)
