"""Host-side skill helpers used outside the rollout (reference: utils/skill_utils.py).

Inside MPPI.command() the suction model and the quaternion costs run in the CUDA kernels. What remains on the host
is what scripts/sim.py calls on the K=1 "real world" env every tick: the suction condition and force
(skill_utils.py:36-94) and the wall-clock pacing helper (:25-33).
"""
import time

import torch


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
