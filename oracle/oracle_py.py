"""ctypes wrapper of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The struct types are the interface types of include/m3p2i_b200.h (m3p2i_b200._abi, types only).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "m3p2i-aip_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from m3p2i_b200 import _abi as A  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("m3p2i_oracle.c", "m3p2i_oracle.h", "point_env.h", "panda_env.h", "halton_spline.h")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "m3p2i_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, fp, ip = A.vp, A.fp, A.ip
        protos = {
            "orc_create": (vp, [C.POINTER(A.Config)]),
            "orc_destroy": (None, [vp]),
            "orc_set_threads": (None, [C.c_int]),
            "orc_get_threads": (C.c_int, []),
            "orc_panda_fk": (C.c_int, [C.POINTER(A.PandaScene), fp, fp, fp]),
            "orc_halton_spline_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                                  C.POINTER(C.c_uint16), C.c_int, fp]),
            "orc_bspline_samples": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, C.c_double, fp]),
        }
        for name, (res, args) in protos.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        skip = ("m3p2i_create", "m3p2i_destroy", "m3p2i_last_error", "m3p2i_version", "m3p2i_device_count")
        L.fn = A.bind(L, prefix_to="orc_", skip=skip)
        _LIB = L
    return _LIB


def set_threads(n):
    lib().orc_set_threads(int(n))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    """Same method surface as m3p2i_b200.native.NativePlanner, computed on the CPU."""

    @classmethod
    def for_sim(cls, sim, cfg=None, noise_mode=A.NOISE_TABLE, seed=0):
        """Backend factory for the sim facade (tests drive the reference's own planner code through it)."""
        from m3p2i_b200 import scene as S
        if cfg is None:
            cfg = S.sim_only_cfg(sim.env_type, sim.num_envs, sim.cfg)
        o = cls(S.build_config(cfg, noise_mode=noise_mode, seed=seed), sim.scene)
        if getattr(cfg.mppi, "filter_u", False) and cfg.mppi.horizon >= 9:
            o.set_filter_matrix(S.savgol_matrix(int(cfg.mppi.horizon)))
        return o

    def __init__(self, config, scene):
        self.L = lib()
        self.cfg = config
        self.h = self.L.orc_create(C.byref(config))
        if not self.h:
            raise ValueError("orc_create: bad config")
        self.K, self.T, self.nu = config.num_samples, config.horizon, config.nu
        self.Kg = config.num_samples_global or config.num_samples
        self.scene = scene
        name = "m3p2i_set_scene_point" if config.env_type == A.ENV_POINT else "m3p2i_set_scene_panda"
        self._ck(self.L.fn[name](self.h, C.byref(scene)), name)
        self.n_actors = scene.n_actors
        self.ndof = 2 if config.env_type == A.ENV_POINT else 9

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle {what} failed rc={rc}")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, dof, root):
        d, r = _f32(dof).ravel(), _f32(root).ravel()
        assert d.size == 2 * self.ndof and r.size == 13 * self.n_actors
        self._ck(self.L.fn["m3p2i_set_state"](self.h, A.as_fp(d), A.as_fp(r)), "set_state")

    def set_objective(self, task, goal, gripper=None):
        g = _f32(goal).ravel()
        self._ck(self.L.fn["m3p2i_set_objective"](self.h, A.TASK_IDS[task], A.as_fp(g), g.size,
                                                 A.GRIPPER_IDS[gripper]), "set_objective")

    def set_noise_table(self, delta):
        d = None if delta is None else _f32(delta)
        if d is not None:
            assert d.shape == (self.K, self.T, self.nu)
        self._ck(self.L.fn["m3p2i_set_noise_table"](self.h, A.as_fp(d)), "set_noise_table")

    def set_noise_halton_spline(self, knot_scale=4, degree=2, smoothing=0.5, perms=None):
        pp, stride = None, 0
        if perms is not None:
            perms = np.ascontiguousarray(perms, np.uint16)
            pp, stride = perms.ctypes.data_as(C.POINTER(C.c_uint16)), int(perms.shape[1])
        self._ck(self.L.fn["m3p2i_set_noise_halton_spline"](self.h, int(knot_scale), int(degree), float(smoothing), pp, stride),
                 "set_noise_halton_spline")

    def set_noise_row0(self, row0):
        d = None if row0 is None else _f32(row0)
        self._ck(self.L.fn["m3p2i_set_noise_row0"](self.h, A.as_fp(d)), "set_noise_row0")

    def get_noise(self):
        out = np.empty((self.K, self.T, self.nu), np.float32)
        self._ck(self.L.fn["m3p2i_get_noise"](self.h, A.as_fp(out)), "get_noise")
        return out

    def get_planner_state(self):
        st = A.PlannerState()
        self._ck(self.L.fn["m3p2i_get_planner_state"](self.h, C.byref(st)), "get_planner_state")
        return st

    def set_planner_state(self, st):
        self._ck(self.L.fn["m3p2i_set_planner_state"](self.h, C.byref(st)), "set_planner_state")

    def set_filter_matrix(self, S):
        s = None if S is None else _f32(S)
        self._ck(self.L.fn["m3p2i_set_filter_matrix"](self.h, A.as_fp(s)), "set_filter_matrix")

    def fetch_result(self, want_cost=True):
        return None, self._last_cost

    def command(self, want_cost=True):
        act = np.empty((self.T, self.nu), np.float32)
        cost = np.empty(self.K, np.float32)
        info = A.CommandInfo()
        self._ck(self.L.fn["m3p2i_command"](self.h, A.as_fp(act), A.as_fp(cost), C.byref(info)), "command")
        self._last_cost = cost
        return act, cost, info

    def rollout_actions(self, actions):
        a = _f32(actions)
        assert a.shape == (self.K, self.T, self.nu)
        st = np.empty((self.K, self.T, 4), np.float32)
        ch = np.empty((self.K, self.T), np.float32)
        self._ck(self.L.fn["m3p2i_rollout_actions"](self.h, A.as_fp(a), A.as_fp(st), A.as_fp(ch)), "rollout_actions")
        return st, ch

    def update_only(self, cost_horizon, actions):
        ch, a = _f32(cost_horizon), _f32(actions)
        out = np.empty((self.T, self.nu), np.float32)
        info = A.CommandInfo()
        self._ck(self.L.fn["m3p2i_update_only"](self.h, A.as_fp(ch), A.as_fp(a), A.as_fp(out), C.byref(info)),
                 "update_only")
        return out, info

    def top_trajs(self, n=A.TOP_N):
        idx = np.empty(n, np.int32)
        w = np.empty(n, np.float32)
        tr = np.empty((n, self.T, 2), np.float32)
        self._ck(self.L.fn["m3p2i_top_trajs"](self.h, n, idx.ctypes.data_as(A.ip), A.as_fp(w), A.as_fp(tr)), "top_trajs")
        return idx, w, tr

    _SHAPES = {A.BUF_ACTIONS: lambda s: (s.K, s.T, s.nu), A.BUF_STATES: lambda s: (s.K, s.T, 4),
               A.BUF_COST_HORIZON: lambda s: (s.K, s.T), A.BUF_COST_DISC: lambda s: (s.Kg,),
               A.BUF_COST_SUM: lambda s: (s.K,), A.BUF_WEIGHTS: lambda s: (3, s.Kg)}

    def read_buffer(self, which):
        out = np.empty(self._SHAPES[which](self), np.float32)
        self._ck(self.L.fn["m3p2i_read_buffer"](self.h, which, A.as_fp(out), out.size), "read_buffer")
        return out

    # multi-rank phases
    def partials_len(self):
        return self.L.fn["m3p2i_partials_len"](self.h)

    def phase_rollout(self):
        J = np.empty(self.K, np.float32)
        self._ck(self.L.fn["m3p2i_phase_rollout"](self.h, A.as_fp(J)), "phase_rollout")
        return J

    def phase_partials(self, J_global):
        J = _f32(J_global)
        assert J.size == self.Kg
        out = np.empty(self.partials_len(), np.float32)
        self._ck(self.L.fn["m3p2i_phase_partials"](self.h, A.as_fp(J), A.as_fp(out)), "phase_partials")
        return out

    def phase_finish(self, partials_sum, want_cost=True):
        p = _f32(partials_sum)
        act = np.empty((self.T, self.nu), np.float32)
        cost = np.empty(self.K, np.float32)
        info = A.CommandInfo()
        self._ck(self.L.fn["m3p2i_phase_finish"](self.h, A.as_fp(p), A.as_fp(act), A.as_fp(cost), C.byref(info)),
                 "phase_finish")
        return act, cost, info

    # persistent K-env sim facade
    def sim_reset(self):
        self._ck(self.L.fn["m3p2i_sim_reset"](self.h), "sim_reset")

    def sim_set_velocity_target(self, u):
        a = _f32(u)
        assert a.shape == (self.K, self.nu)
        self._ck(self.L.fn["m3p2i_sim_set_velocity_target"](self.h, A.as_fp(a)), "sim_set_velocity_target")

    def sim_apply_forces(self, f_robot, f_box):
        fr = None if f_robot is None else _f32(f_robot)
        fb = None if f_box is None else _f32(f_box)
        self._ck(self.L.fn["m3p2i_sim_apply_forces"](self.h, A.as_fp(fr), A.as_fp(fb)), "sim_apply_forces")

    def sim_write(self, dof, root):
        d = None if dof is None else _f32(dof)
        r = None if root is None else _f32(root)
        self._ck(self.L.fn["m3p2i_sim_write"](self.h, A.as_fp(d), A.as_fp(r)), "sim_write")

    def sim_cost(self):
        out = np.empty(self.K, np.float32)
        self._ck(self.L.fn["m3p2i_sim_cost"](self.h, A.as_fp(out)), "sim_cost")
        return out

    def sim_step(self):
        self._ck(self.L.fn["m3p2i_sim_step"](self.h), "sim_step")

    def sim_read(self):
        n_link = 1 if self.cfg.env_type == A.ENV_POINT else 3
        n_con = 1 if self.cfg.env_type == A.ENV_POINT else 3
        dof = np.empty((self.K, 2 * self.ndof), np.float32)
        root = np.empty((self.K, self.n_actors, 13), np.float32)
        link = np.empty((self.K, n_link, 13), np.float32)
        con = np.empty((self.K, n_con, 3), np.float32)
        self._ck(self.L.fn["m3p2i_sim_read"](self.h, A.as_fp(dof), A.as_fp(root), A.as_fp(link), A.as_fp(con)), "sim_read")
        return dof, root, link, con


def panda_fk(scene, q, qd=None):
    L = lib()
    out = np.empty((3, 13), np.float32)
    qq = _f32(q)
    qv = None if qd is None else _f32(qd)
    rc = L.orc_panda_fk(C.byref(scene), A.as_fp(qq), A.as_fp(qv), A.as_fp(out))
    if rc:
        raise RuntimeError("orc_panda_fk failed")
    return out


def bspline_samples(cv, T, degree=2, smoothing=0.5):
    """skill_utils.bspline (scipy splrep s=0.5 + splev) restated in C (halton_spline.h)."""
    cv = _f32(cv)
    out = np.empty(T, np.float32)
    if lib().orc_bspline_samples(A.as_fp(cv), len(cv), int(T), int(degree), float(smoothing), A.as_fp(out)):
        raise RuntimeError("orc_bspline_samples failed")
    return out


def halton_spline_table(K, T, nu, knot_scale=4, degree=2, smoothing=0.5, offset=0, perms=None):
    """delta [K, T, nu] of the halton-spline mode (mppi.py:458-478) for global samples offset .. offset + K.
    perms: None (plain Halton) or uint16 [n_knots * nu, stride] digit permutations (generalised Halton)."""
    out = np.empty((K, T, nu), np.float32)
    pp, stride = None, 0
    if perms is not None:
        perms = np.ascontiguousarray(perms, np.uint16)
        pp, stride = perms.ctypes.data_as(C.POINTER(C.c_uint16)), perms.shape[1]
    if lib().orc_halton_spline_table(int(K), int(offset), int(T), int(nu), int(knot_scale), int(degree), float(smoothing),
                                     pp, stride, A.as_fp(out)):
        raise RuntimeError("orc_halton_spline_table failed")
    return out
