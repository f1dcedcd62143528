"""ctypes mirror of include/m3p2i_b200.h (types and constants only).

The structures here must stay byte-compatible with the C header; tests/test_abi.py checks sizeof() of every
struct against the values the compiled library reports (m3p2i_abi_sizeof).
"""
import ctypes as C

MAX_NU = 9
MAX_STATIC = 8
NX = 4
TOP_N = 20
MAX_HORIZON = 64

ENV_POINT, ENV_PANDA = 0, 1
ENV_IDS = {"point_env": ENV_POINT, "panda_env": ENV_PANDA}

TASK_IDS = {"navigation": 0, "push": 1, "pull": 2, "push_pull": 3, "reach": 4, "pick": 5, "place": 6}
GRIPPER_IDS = {None: 0, "none": 0, "open": 1, "close": 2}

NOISE_TABLE, NOISE_PHILOX, NOISE_PHILOX_SPLINE = 0, 1, 2

BUF_ACTIONS, BUF_STATES, BUF_COST_HORIZON, BUF_COST_DISC, BUF_COST_SUM, BUF_WEIGHTS, BUF_NOISE = range(7)

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_STATE", -4: "ERR_NCCL", -5: "ERR_NO_DEVICE"}

f32 = C.c_float
i32 = C.c_int32


class Config(C.Structure):
    _fields_ = [
        ("env_type", i32), ("num_samples", i32), ("horizon", i32), ("nu", i32), ("multi_modal", i32),
        ("sample_null_action", i32), ("filter_u", i32), ("noise_mode", i32), ("num_samples_global", i32),
        ("sample_offset", i32), ("substeps", i32), ("solver_passes", i32), ("lanes_per_sample", i32),
        ("update_cov", i32), ("reserved_i", i32 * 2),
        ("dt", f32), ("gamma", f32), ("step_size_mean", f32), ("u_scale", f32), ("kp_suction", f32),
        ("pre_height_diff", f32), ("tilt_cos_theta", f32), ("reserved_f", f32),
        ("u_min", f32 * MAX_NU), ("u_max", f32 * MAX_NU), ("sigma", f32 * MAX_NU),
        ("seed", C.c_uint64),
    ]


class Box(C.Structure):
    _fields_ = [("pos", f32 * 3), ("half", f32 * 3), ("quat", f32 * 4), ("mu", f32), ("actor", i32)]


class Body(C.Structure):
    _fields_ = [("half", f32 * 3), ("mass", f32), ("inertia", f32), ("mu", f32), ("r_eff", f32), ("actor", i32)]


class PointScene(C.Structure):
    _fields_ = [
        ("robot_radius", f32), ("robot_mass", f32), ("robot_mu", f32), ("drive_damping", f32), ("drive_effort", f32),
        ("gravity", f32), ("ground_mu", f32), ("contact_margin", f32), ("baumgarte", f32), ("slop", f32),
        ("max_corr_vel", f32),
        ("n_static", i32), ("n_actors", i32), ("reserved", i32),
        ("box", Body), ("dyn_obs", Body), ("statics", Box * MAX_STATIC),
    ]


class PandaScene(C.Structure):
    _fields_ = [
        ("base_pos", f32 * 3), ("gravity", f32),
        ("q_lower", f32 * MAX_NU), ("q_upper", f32 * MAX_NU), ("qd_limit", f32 * MAX_NU), ("effort", f32 * MAX_NU),
        ("drive_damping", f32), ("arm_inertia", f32), ("finger_mass", f32), ("robot_mu", f32),
        ("finger_half", f32 * 3), ("finger_center", f32 * 3), ("hand_half", f32 * 3), ("hand_center", f32 * 3),
        ("contact_margin", f32), ("baumgarte", f32), ("slop", f32), ("max_corr_vel", f32), ("penalty_stiffness", f32),
        ("joint_inertia", f32 * MAX_NU), ("warm_start", f32), ("sleep_lin", f32), ("sleep_ang", f32), ("sleep_gap", f32),
        ("n_static", i32), ("n_actors", i32), ("idx_table", i32), ("idx_shelf", i32), ("link_sweeps", i32), ("report_cube_contacts", i32),
        ("cube_a", Body), ("cube_b", Body), ("statics", Box * MAX_STATIC),
    ]


class PlannerState(C.Structure):
    _fields_ = [
        ("mean_action", f32 * (MAX_HORIZON * MAX_NU)), ("mean_action_1", f32 * (MAX_HORIZON * MAX_NU)),
        ("mean_action_2", f32 * (MAX_HORIZON * MAX_NU)), ("best_traj", f32 * (MAX_HORIZON * MAX_NU)),
        ("best_traj_1", f32 * (MAX_HORIZON * MAX_NU)), ("best_traj_2", f32 * (MAX_HORIZON * MAX_NU)),
        ("beta", C.c_double), ("cov_action", f32 * MAX_NU), ("reserved", f32),
    ]


class CommandInfo(C.Structure):
    _fields_ = [
        ("eta", f32 * 3), ("beta", f32 * 3), ("min_cost", f32 * 3), ("best_idx", i32 * 3),
        ("weight_push", f32), ("weight_pull", f32), ("mean_cost_sum", f32), ("kernel_ms", f32),
        ("launches", i32), ("beta_iters", i32), ("rollout_ms", f32), ("rollout_lanes", i32), ("peer_wait_ms", f32 * 2), ("near_samples", i32),
    ]


class PeerHandle(C.Structure):
    _fields_ = [("ipc", C.c_ubyte * 64), ("pid", C.c_int64), ("ptr", C.c_uint64), ("bytes", C.c_uint64),
                ("device", i32), ("reserved", i32)]


STRUCTS = {"M3P2IPeerHandle": PeerHandle, "M3P2IConfig": Config, "M3P2IBox": Box, "M3P2IBody": Body, "M3P2IPointScene": PointScene,
           "M3P2IPandaScene": PandaScene, "M3P2IPlannerState": PlannerState, "M3P2ICommandInfo": CommandInfo}

fp = C.POINTER(f32)
ip = C.POINTER(i32)
vp = C.c_void_p

# name -> (restype, argtypes) for every entry point declared in include/m3p2i_b200.h (handle passed as void*)
PROTOTYPES = {
    "m3p2i_last_error": (C.c_char_p, []),
    "m3p2i_version": (C.c_int, []),
    "m3p2i_device_count": (C.c_int, []),
    "m3p2i_abi_sizeof": (C.c_int, [C.c_char_p]),
    "m3p2i_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(vp)]),
    "m3p2i_destroy": (None, [vp]),
    "m3p2i_set_scene_point": (C.c_int, [vp, C.POINTER(PointScene)]),
    "m3p2i_set_scene_panda": (C.c_int, [vp, C.POINTER(PandaScene)]),
    "m3p2i_set_state": (C.c_int, [vp, fp, fp]),
    "m3p2i_set_objective": (C.c_int, [vp, C.c_int, fp, C.c_int, C.c_int]),
    "m3p2i_set_noise_table": (C.c_int, [vp, fp]),
    "m3p2i_set_noise_row0": (C.c_int, [vp, fp]),
    "m3p2i_set_noise_halton_spline": (C.c_int, [vp, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_uint16), C.c_int]),
    "m3p2i_get_noise": (C.c_int, [vp, fp]),
    "m3p2i_get_planner_state": (C.c_int, [vp, C.POINTER(PlannerState)]),
    "m3p2i_set_planner_state": (C.c_int, [vp, C.POINTER(PlannerState)]),
    "m3p2i_set_filter_matrix": (C.c_int, [vp, fp]),
    "m3p2i_command": (C.c_int, [vp, fp, fp, C.POINTER(CommandInfo)]),
    "m3p2i_command_resident": (C.c_int, [vp, C.POINTER(CommandInfo)]),
    "m3p2i_fetch_result": (C.c_int, [vp, fp, fp]),
    "m3p2i_rollout_actions": (C.c_int, [vp, fp, fp, fp]),
    "m3p2i_sample_actions": (C.c_int, [vp, fp]),
    "m3p2i_update_only": (C.c_int, [vp, fp, fp, fp, C.POINTER(CommandInfo)]),
    "m3p2i_top_trajs": (C.c_int, [vp, C.c_int, ip, fp, fp]),
    "m3p2i_get_buffer": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "m3p2i_read_buffer": (C.c_int, [vp, C.c_int, fp, C.c_size_t]),
    "m3p2i_sim_reset": (C.c_int, [vp]),
    "m3p2i_sim_set_velocity_target": (C.c_int, [vp, fp]),
    "m3p2i_sim_apply_forces": (C.c_int, [vp, fp, fp]),
    "m3p2i_sim_step": (C.c_int, [vp]),
    "m3p2i_sim_write": (C.c_int, [vp, fp, fp]),
    "m3p2i_sim_cost": (C.c_int, [vp, fp]),
    "m3p2i_sim_read": (C.c_int, [vp, fp, fp, fp, fp]),
    "m3p2i_partials_len": (C.c_int, [vp]),
    "m3p2i_phase_rollout": (C.c_int, [vp, fp]),
    "m3p2i_phase_partials": (C.c_int, [vp, fp, fp]),
    "m3p2i_phase_finish": (C.c_int, [vp, fp, fp, fp, C.POINTER(CommandInfo)]),
    "m3p2i_comm_unique_id": (C.c_int, [vp]),
    "m3p2i_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "m3p2i_peer_export": (C.c_int, [vp, C.POINTER(PeerHandle)]),
    "m3p2i_peer_attach": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(PeerHandle)]),
    "m3p2i_set_stream": (C.c_int, [vp, vp]),
}


def bind(lib, prototypes=PROTOTYPES, prefix_from="m3p2i_", prefix_to=None, skip=()):
    """Attach restype/argtypes. With prefix_to, binds the same signatures under another symbol prefix (used by
    test doubles that export the same argument lists)."""
    out = {}
    for name, (res, args) in prototypes.items():
        if name in skip:
            continue
        sym = name if prefix_to is None else prefix_to + name[len(prefix_from):]
        try:
            fn = getattr(lib, sym)
        except AttributeError:
            continue
        fn.restype = res
        fn.argtypes = args
        out[name] = fn
    return out


def as_fp(a):
    """float32 C-contiguous numpy array -> float* (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(fp)
