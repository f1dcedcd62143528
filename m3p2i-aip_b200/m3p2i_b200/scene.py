"""Scene description: the reference's per-actor YAMLs -> the flat structs the rollout kernels consume.

Restates what actor_utils.load_env_cfgs + IsaacGymWrapper.creat_env/_create_actor do for the two environments the
reference ships (reference: utils/isaacgym_utils/actor_utils.py:16-46,94-101; isaacgym_wrapper.py:242-352), with
an explicit, deterministic actor order (the reference iterates Path.iterdir() unsorted, actor_utils.py:97):
non-robot actors sorted by the numeric prefix of their file name, robots last. calculate_suction relies on the
robot body being last (skill_utils.py:89-90).

`default_actors(env_type)` holds the shipped scene constants (config/point_env/*.yaml, config/panda_env/*.yaml,
assets/urdf/pointRobot.urdf, assets/urdf/franka_description/robots/franka_panda.urdf); `load_actor_dir` reads a
directory of reference-style YAMLs instead.
"""
import math
import os
import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi as A

BOX_DENSITY = 1000.0  # IsaacGym default asset density; `mass` in the YAMLs is never applied (isaacgym_wrapper.py:293-300)


@dataclass
class Actor:
    """Subset of actor_utils.ActorWrapper that the integrator needs (same field names and defaults)."""
    type: str
    name: str
    init_pos: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    init_pos_on_table: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    init_pos_on_shelf: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    init_ori: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 1.0])
    size: List[float] = field(default_factory=lambda: [0.1, 0.1, 0.1])
    fixed: bool = False
    collision: bool = True
    friction: float = 1.0
    gravity: bool = True
    urdf_file: Optional[str] = None
    init_joint_pose: Optional[List[float]] = None


def _point_actors():
    s2 = 0.707107
    return [
        Actor("box", "wall-1", init_pos=[4.0, 0.0, 0.0], size=[0.1, 8.0, 0.2], fixed=True),
        Actor("box", "wall-2", init_pos=[-4.0, 0.0, 0.0], size=[0.1, 8.0, 0.2], fixed=True),
        Actor("box", "wall-3", init_pos=[0.0, 4.0, 0.0], init_ori=[0.0, 0.0, s2, s2], size=[0.1, 8.0, 0.2], fixed=True),
        Actor("box", "wall-4", init_pos=[0.0, -4.0, 0.0], init_ori=[0.0, 0.0, s2, s2], size=[0.1, 8.0, 0.2], fixed=True),
        Actor("box", "obs", init_pos=[2.0, 2.0, 0.0], size=[0.3, 0.4, 0.5], fixed=True),
        Actor("box", "dyn-obs", init_pos=[-2.0, 2.0, 0.0], size=[0.4, 0.4, 0.1]),
        Actor("box", "box", init_pos=[0.0, 2.0, 0.0], size=[0.4, 0.4, 0.1], friction=0.5),
        Actor("box", "goal", init_pos=[-3.75, -3.75, 0.0], size=[0.45, 0.45, 0.01], fixed=True, collision=False),
        Actor("box", "yaxis", init_pos=[0.0, 0.25, 0.01], size=[0.05, 0.5, 0.01], fixed=True, collision=False),
        Actor("box", "xaxis", init_pos=[0.25, 0.0, 0.01], size=[0.5, 0.05, 0.01], fixed=True, collision=False),
        Actor("robot", "point_robot", init_pos=[0.0, 0.0, 0.05], fixed=True, friction=0.05, urdf_file="pointRobot.urdf"),
    ]


def _panda_actors():
    return [
        Actor("box", "table", init_pos=[0.0, 0.0, 1.0], size=[1.2, 1.2, 0.05], fixed=True),
        Actor("box", "table_stand", init_pos=[-0.5, 0.0, 1.075], size=[0.2, 0.2, 0.1], fixed=True),
        Actor("box", "shelf_stand", init_pos=[0.5, 0.0, 1.175], size=[0.2, 0.2, 0.3], fixed=True),
        Actor("box", "dyn-obs", init_pos=[0.35, 0.0, 1.735], size=[0.2, 0.2, 0.02], gravity=False),
        Actor("box", "cubeA", init_pos_on_table=[0.2, -0.2, 1.06], init_pos_on_shelf=[0.425, 0.0, 1.35],
              size=[0.05, 0.05, 0.05]),
        Actor("box", "cubeB", init_pos=[0.2, 0.2, 1.06], size=[0.05, 0.05, 0.05]),
        Actor("robot", "panda", init_pos=[-0.45, 0.0, 1.125], fixed=True, gravity=False,
              urdf_file="franka_description/robots/franka_panda.urdf",
              init_joint_pose=[0, 0, 0, 0, 0, 0, -2, 0, 0, 0, 1.8675, 0, 0, 0, 0.02, 0, 0.02, 0]),
    ]


def default_actors(env_type):
    if env_type == "point_env":
        return _point_actors()
    if env_type == "panda_env":
        return _panda_actors()
    raise ValueError(f"unknown env_type {env_type!r}")


def load_actor_dir(path):
    """Reference-style directory of per-actor YAML files -> ordered actor list."""
    import yaml
    entries = []
    for fn in os.listdir(path):
        if not fn.endswith((".yaml", ".yml")):
            continue
        with open(os.path.join(path, fn)) as f:
            d = yaml.safe_load(f)
        known = {k: v for k, v in d.items() if k in Actor.__dataclass_fields__}
        m = re.match(r"(\d+)_", fn)
        entries.append((1 if d.get("type") == "robot" else 0, int(m.group(1)) if m else 10 ** 6, fn, Actor(**known)))
    entries.sort(key=lambda e: e[:3])
    return [e[3] for e in entries]


def actor_index(actors, name):
    return [a.name for a in actors].index(name)


def initial_root_state(actors, cube_on_shelf=False):
    """[n_actors, 13] rows pos3, quat4 xyzw, linvel3, angvel3 (isaacgym_wrapper.py:269-283)."""
    root = np.zeros((len(actors), 13), np.float32)
    for i, a in enumerate(actors):
        pos = a.init_pos
        if a.name == "cubeA":
            pos = a.init_pos_on_shelf if cube_on_shelf else a.init_pos_on_table
        root[i, 0:3] = pos
        root[i, 3:7] = a.init_ori
    return root


def initial_dof_state(actors):
    """[2*ndof] interleaved (pos, vel) (isaacgym_wrapper.py:222-240)."""
    robot = [a for a in actors if a.type == "robot"][-1]
    if robot.init_joint_pose:
        return np.asarray(robot.init_joint_pose, np.float32)
    ndof = 2 if "pointRobot" in (robot.urdf_file or "") else 9
    return np.zeros(2 * ndof, np.float32)


def _fill(arr, values):
    for i, v in enumerate(values):
        arr[i] = float(v)


def _static_box(a, actor_row=-1):
    b = A.Box()
    _fill(b.pos, a.init_pos)
    _fill(b.half, [0.5 * s for s in a.size])
    _fill(b.quat, a.init_ori)
    b.mu = a.friction
    b.actor = actor_row
    return b


def _planar_body(a, row):
    sx, sy, sz = a.size
    m = BOX_DENSITY * sx * sy * sz
    b = A.Body()
    _fill(b.half, [0.5 * sx, 0.5 * sy, 0.5 * sz])
    b.mass = m
    b.inertia = m * (sx * sx + sy * sy) / 12.0
    b.mu = a.friction
    # mean distance of a uniformly loaded rectangle's area from its centre (square: 0.3826 * side)
    b.r_eff = 0.3826 * 0.5 * (sx + sy)
    b.actor = row
    return b


def _cube_body(a, row):
    sx, sy, sz = a.size
    m = BOX_DENSITY * sx * sy * sz
    b = A.Body()
    _fill(b.half, [0.5 * sx, 0.5 * sy, 0.5 * sz])
    b.mass = m
    b.inertia = m * (sx * sx + sy * sy + sz * sz) / 18.0  # isotropic: mean of the three principal moments
    b.mu = a.friction
    b.r_eff = 0.0
    b.actor = row
    return b


def build_point_scene(actors=None):
    actors = actors or _point_actors()
    robot = [a for a in actors if a.type == "robot"][-1]
    s = A.PointScene()
    s.robot_radius = 0.2      # pointRobot.urdf:17
    s.robot_mass = 10.0       # pointRobot.urdf:11
    s.robot_mu = robot.friction
    s.drive_damping = 600.0   # isaacgym_wrapper.py:344
    s.drive_effort = 1000.0   # pointRobot.urdf:36
    s.gravity = 9.8           # isaacgym_wrapper.py:25
    s.ground_mu = 1.0         # isaacgym_wrapper.py:466
    s.contact_margin = 0.01   # isaacgym_wrapper.py:30
    s.baumgarte = 0.2
    s.slop = 0.002
    s.max_corr_vel = 2.0
    s.n_actors = len(actors)
    n = 0
    for i, a in enumerate(actors):
        if a.type != "box" or not a.collision:
            continue
        if a.name == "box":
            s.box = _planar_body(a, i)
        elif a.name == "dyn-obs":
            s.dyn_obs = _planar_body(a, i)
        elif a.fixed:
            if n >= A.MAX_STATIC:
                raise ValueError("too many static boxes")
            s.statics[n] = _static_box(a)
            n += 1
    s.n_static = n
    return s


# field -> value applied to every M3P2IPandaScene built afterwards (solver experiments, tests that need the physical
# table forces: {"report_cube_contacts": 1}, {"link_sweeps": 1}, ...)
PANDA_SCENE_OVERRIDES = {}


def build_panda_scene(actors=None):
    actors = actors or _panda_actors()
    robot = [a for a in actors if a.type == "robot"][-1]
    s = A.PandaScene()
    _fill(s.base_pos, robot.init_pos)
    s.gravity = 9.8
    # franka_panda.urdf <limit> of joints 1-7 and the two finger joints
    _fill(s.q_lower, [-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973, 0.0, 0.0])
    _fill(s.q_upper, [2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973, 0.04, 0.04])
    _fill(s.qd_limit, [2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61, 0.2, 0.2])
    _fill(s.effort, [87, 87, 87, 87, 12, 12, 12, 20, 20])
    s.drive_damping = 600.0
    s.arm_inertia = 0.1
    s.finger_mass = 0.1
    s.robot_mu = robot.friction
    # AABBs of meshes/collision/{finger,hand}.obj in the link frames
    _fill(s.finger_half, [0.0105, 0.0132, 0.0269])
    _fill(s.finger_center, [0.0, 0.0132, 0.0269])
    _fill(s.hand_half, [0.0316, 0.1022, 0.04595])
    _fill(s.hand_center, [0.0, -0.0018, 0.02005])
    s.contact_margin = 0.01
    s.baumgarte = 0.2
    s.slop = 0.0005
    s.max_corr_vel = 0.5
    s.penalty_stiffness = 2000.0
    # link masses from the collision meshes at IsaacGym's default density, reflected onto the joint axes at the
    # initial pose (tools/panda_inertia.py); the fingers keep finger_mass
    _fill(s.joint_inertia, [1.3195, 2.1231, 1.2994, 0.9183, 0.0271, 0.0366, 0.0030, 0.0, 0.0])
    s.warm_start = 0.9
    s.sleep_lin, s.sleep_ang, s.sleep_gap = 5e-3, 5e-2, 2e-3
    s.link_sweeps = 2
    s.report_cube_contacts = 0
    s.n_actors = len(actors)
    s.idx_table = -1
    s.idx_shelf = -1
    n = 0
    for i, a in enumerate(actors):
        if a.type != "box" or not a.collision:
            continue
        if a.name == "cubeA":
            s.cube_a = _cube_body(a, i)
        elif a.name == "cubeB":
            s.cube_b = _cube_body(a, i)
        else:
            if n >= A.MAX_STATIC:
                raise ValueError("too many static boxes")
            # movable boxes without gravity (the floating dyn-obs plate) are held where the real state puts them
            s.statics[n] = _static_box(a, -1 if a.fixed else i)
            if a.name == "table":
                s.idx_table = n
            if a.name == "shelf_stand":
                s.idx_shelf = n
            n += 1
    # the ground plane of the reference (z = 0, friction 1, isaacgym_wrapper.py:462-469) as a thick slab: a cube that
    # leaves the table lands on it instead of falling for ever
    if n < A.MAX_STATIC:
        g = A.Box()
        _fill(g.pos, [0.0, 0.0, -0.5])
        _fill(g.half, [10.0, 10.0, 0.5])
        _fill(g.quat, [0.0, 0.0, 0.0, 1.0])
        g.mu = 1.0
        g.actor = -1
        s.statics[n] = g
        n += 1
    s.n_static = n
    for key, value in PANDA_SCENE_OVERRIDES.items():
        setattr(s, key, value)
    return s


def build_config(cfg, num_samples_local=None, sample_offset=0, noise_mode=A.NOISE_TABLE, seed=0):
    """Attribute-style reference cfg (ExampleConfig / hydra DictConfig / SimpleNamespace) -> M3P2IConfig."""
    m = cfg.mppi
    c = A.Config()
    c.env_type = A.ENV_IDS[cfg.env_type]
    K = int(m.num_samples)
    c.num_samples_global = K
    c.num_samples = K if num_samples_local is None else int(num_samples_local)
    c.sample_offset = int(sample_offset)
    c.horizon = int(m.horizon)
    sig = np.asarray(m.noise_sigma, np.float64)
    nu = sig.shape[0]
    c.nu = nu
    c.multi_modal = int(bool(cfg.multi_modal))
    c.sample_null_action = int(bool(getattr(m, "sample_null_action", False)))
    c.filter_u = int(bool(getattr(m, "filter_u", False)))
    c.noise_mode = noise_mode
    c.update_cov = int(bool(getattr(m, "update_cov", False)))
    ig = getattr(cfg, "isaacgym", None)
    c.substeps = int(getattr(ig, "substeps", 2)) if ig is not None else 2
    c.solver_passes = 2
    c.lanes_per_sample = int(getattr(m, "lanes_per_sample", 0))  # 0 = let the library choose
    c.dt = float(getattr(ig, "dt", 0.05 if c.env_type == A.ENV_POINT else 0.01)) if ig is not None else \
        (0.05 if c.env_type == A.ENV_POINT else 0.01)
    c.gamma = float(getattr(m, "rollout_var_discount", 0.95))
    c.step_size_mean = 0.98  # mppi.py:178
    c.u_scale = float(getattr(m, "u_scale", 1.0))
    c.kp_suction = float(getattr(cfg, "kp_suction", 0.0))
    c.pre_height_diff = float(getattr(cfg, "pre_height_diff", 0.0))
    c.tilt_cos_theta = 0.5   # cost_functions.py:13
    u_min, u_max = getattr(m, "u_min", None), getattr(m, "u_max", None)
    if u_max is not None and u_min is None:
        u_min = [-v for v in u_max]
    if u_min is not None and u_max is None:
        u_max = [-v for v in u_min]
    _fill(c.u_min, u_min)
    _fill(c.u_max, u_max)
    # torch.sqrt(torch.diagonal(noise_sigma)) evaluated in fp32 (mppi.py:175-176)
    _fill(c.sigma, np.sqrt(np.diagonal(sig).astype(np.float32)))
    c.seed = int(seed)
    return c


def savgol_matrix(T, window=9, order=2):
    """The Savitzky-Golay smoothing of mppi.py:257-263 as a fixed [T,T] matrix: filtered = S @ action."""
    from scipy import signal
    if window % 2 == 0:
        window -= 1
    eye = np.eye(T, dtype=np.float64)
    S = signal.savgol_filter(eye, window, order, deriv=0, delta=1.0, axis=0, mode="interp", cval=0.0)
    return np.ascontiguousarray(S, dtype=np.float32)


def yaw_quat(theta):
    return [0.0, 0.0, math.sin(0.5 * theta), math.cos(0.5 * theta)]


def sim_only_cfg(env_type, num_envs, isaacgym_cfg=None):
    """Planner-less configuration for a sim facade that is only stepped (e.g. the K=1 'real world' of sim.py)."""
    from types import SimpleNamespace as NS
    nu = 2 if env_type == "point_env" else 9
    mppi = NS(num_samples=num_envs, horizon=1, noise_sigma=np.eye(nu).tolist(), u_min=[-1e9] * nu, u_max=[1e9] * nu,
              sample_null_action=False, filter_u=False, rollout_var_discount=1.0, u_scale=1.0)
    return NS(env_type=env_type, multi_modal=False, mppi=mppi, isaacgym=isaacgym_cfg, kp_suction=0.0,
              pre_height_diff=0.0)


def make_cfg(env_type="point_env", task="navigation", goal=None, num_samples=200, horizon=None, multi_modal=False,
             cube_on_shelf=False, device="cpu", **mppi_overrides):
    """The shipped hydra configuration tree (config/config_{point,panda}.yaml + mppi/*.yaml + isaacgym/*.yaml)
    as a plain attribute-style object."""
    from types import SimpleNamespace as NS
    if env_type == "point_env":
        mppi = dict(num_samples=num_samples, horizon=horizon or 15, nx=4, lambda_=0.5, u_min=[-3.0, -3.0],
                    u_max=[3.0, 3.0], noise_sigma=[[3.0, 0.0], [0.0, 3.0]], u_per_command=horizon or 15)
        ig = NS(dt=0.05, substeps=2)
        top = dict(task=task, goal=goal if goal is not None else [-3.75, -3.75], kp_suction=400, suction_active=True,
                   pre_height_diff=0.0)
    else:
        sig = np.diag([10.0] * 7 + [0.8] * 2).tolist()
        mppi = dict(num_samples=num_samples, horizon=horizon or 12, nx=18, lambda_=0.05,
                    u_min=[-2.0] * 7 + [-1.5] * 2, u_max=[2.0] * 7 + [1.5] * 2, noise_sigma=sig,
                    u_per_command=horizon or 12)
        ig = NS(dt=0.01, substeps=2)
        top = dict(task=task, goal=goal if goal is not None else [0.0] * 7, kp_suction=0, suction_active=False,
                   pre_height_diff=0.05)
    base = dict(mppi_mode="halton-spline", sampling_method="halton", device=device, noise_mu=None, U_init=None,
                u_init=0.0, u_scale=1, rollout_var_discount=0.95, sample_null_action=True, filter_u=True,
                use_priors=False, update_cov=False, update_lambda=False, noise_abs_cost=False, seed_val=0)
    base.update(mppi)
    base.update(mppi_overrides)
    return NS(env_type=env_type, multi_modal=multi_modal, cube_on_shelf=cube_on_shelf, mppi=NS(**base), isaacgym=ig,
              **top)
