"""ctypes binding of libm3p2i_b200.so, the B200-native rollout + update hot path.

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, construction fails.
Build the library with `python -m m3p2i_b200.build` (or __graft_entry__.build()).
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A
from . import scene as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm3p2i_b200.so")
_LIB = None


class NativeError(RuntimeError):
    pass


def lib():
    """Loads the C-ABI library once; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: build it with `python -m m3p2i_b200.build` "
                              "(the product has no CPU path)")
        L = C.CDLL(LIB_PATH)
        L.fn = A.bind(L)
        missing = [n for n in A.PROTOTYPES if n not in L.fn]
        if missing:
            raise NativeError(f"libm3p2i_b200.so does not export {missing}")
        for name, struct in A.STRUCTS.items():
            got = L.fn["m3p2i_abi_sizeof"](name.encode())
            if got != C.sizeof(struct):
                raise NativeError(f"ABI mismatch: sizeof({name}) is {got} in the library, {C.sizeof(struct)} in _abi.py")
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class NativePlanner:
    """One handle = K rollout environments + planner state on one GPU (K may be a shard of a larger batch)."""

    def __init__(self, config, scene, device=0):
        self.L = lib()
        self.fn = self.L.fn
        self.cfg = config
        self.h = A.vp()
        self._ck(self.fn["m3p2i_create"](C.byref(config), int(device), C.byref(self.h)), "m3p2i_create")
        self.K, self.T, self.nu = config.num_samples, config.horizon, config.nu
        self.Kg = config.num_samples_global or config.num_samples
        self.scene = scene
        name = "m3p2i_set_scene_point" if config.env_type == A.ENV_POINT else "m3p2i_set_scene_panda"
        self._ck(self.fn[name](self.h, C.byref(scene)), name)
        self.n_actors = scene.n_actors
        self.ndof = 2 if config.env_type == A.ENV_POINT else 9
        self._act = np.empty((self.T, self.nu), np.float32)
        self._cost = np.empty(self.K, np.float32)
        self._info = A.CommandInfo()

    @classmethod
    def for_sim(cls, sim, cfg=None, noise_mode=A.NOISE_TABLE, seed=0, device=0):
        if cfg is None:
            cfg = S.sim_only_cfg(sim.env_type, sim.num_envs, sim.cfg)
        p = cls(S.build_config(cfg, noise_mode=noise_mode, seed=seed), sim.scene, device=device)
        if getattr(cfg.mppi, "filter_u", False) and cfg.mppi.horizon >= 9:
            p.set_filter_matrix(S.savgol_matrix(int(cfg.mppi.horizon)))
        return p

    def _ck(self, rc, what):
        if rc != 0:
            msg = self.fn["m3p2i_last_error"]().decode(errors="replace")
            raise NativeError(f"{what} failed ({A.ERR_NAMES.get(rc, rc)}): {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.fn["m3p2i_destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- inputs
    def set_state(self, dof, root):
        d, r = _f32(dof).ravel(), _f32(root).ravel()
        if d.size != 2 * self.ndof or r.size != 13 * self.n_actors:
            raise ValueError(f"dof_state needs {2 * self.ndof} floats and root_state {13 * self.n_actors}")
        self._ck(self.fn["m3p2i_set_state"](self.h, A.as_fp(d), A.as_fp(r)), "m3p2i_set_state")

    def set_objective(self, task, goal, gripper=None):
        g = _f32(goal).ravel()
        self._ck(self.fn["m3p2i_set_objective"](self.h, A.TASK_IDS[task], A.as_fp(g), g.size, A.GRIPPER_IDS[gripper]),
                 "m3p2i_set_objective")

    def set_noise_table(self, delta):
        d = None if delta is None else _f32(delta)
        if d is not None and d.shape != (self.K, self.T, self.nu):
            raise ValueError(f"delta must be [{self.K},{self.T},{self.nu}]")
        self._ck(self.fn["m3p2i_set_noise_table"](self.h, A.as_fp(d)), "m3p2i_set_noise_table")

    def set_noise_halton_spline(self, knot_scale=4, degree=2, smoothing=0.5, perms=None):
        """The reference's once-sampled halton-spline table (mppi.py:458-478) built on the device for this shard's global
        samples. perms: None (plain Halton) or uint16 [n_knots * nu, stride] digit permutations (ghalton.EA_PERMS)."""
        pp, stride = None, 0
        if perms is not None:
            perms = np.ascontiguousarray(perms, np.uint16)
            pp, stride = perms.ctypes.data_as(C.POINTER(C.c_uint16)), int(perms.shape[1])
        self._ck(self.fn["m3p2i_set_noise_halton_spline"](self.h, int(knot_scale), int(degree), float(smoothing), pp, stride),
                 "m3p2i_set_noise_halton_spline")

    def set_noise_row0(self, row0):
        d = None if row0 is None else _f32(row0)
        self._ck(self.fn["m3p2i_set_noise_row0"](self.h, A.as_fp(d)), "m3p2i_set_noise_row0")

    def get_noise(self):
        out = np.empty((self.K, self.T, self.nu), np.float32)
        self._ck(self.fn["m3p2i_get_noise"](self.h, A.as_fp(out)), "m3p2i_get_noise")
        return out

    def get_planner_state(self):
        st = A.PlannerState()
        self._ck(self.fn["m3p2i_get_planner_state"](self.h, C.byref(st)), "m3p2i_get_planner_state")
        return st

    def set_planner_state(self, st):
        self._ck(self.fn["m3p2i_set_planner_state"](self.h, C.byref(st)), "m3p2i_set_planner_state")

    def set_filter_matrix(self, Smat):
        s = None if Smat is None else _f32(Smat)
        self._ck(self.fn["m3p2i_set_filter_matrix"](self.h, A.as_fp(s)), "m3p2i_set_filter_matrix")

    def set_stream(self, cuda_stream):
        self._ck(self.fn["m3p2i_set_stream"](self.h, A.vp(cuda_stream or 0)), "m3p2i_set_stream")

    # ---------------------------------------------------------------- one tick
    def command(self, want_cost=True, want_info=True):
        """-> (action [T,nu], cost_total [K] or None, CommandInfo or None). Arrays are reused between calls.
        want_info=False: no scalars and no timing events on the stream (the plain control loop)."""
        self._ck(self.fn["m3p2i_command"](self.h, A.as_fp(self._act), A.as_fp(self._cost) if want_cost else None,
                                         C.byref(self._info) if want_info else None), "m3p2i_command")
        return self._act, (self._cost if want_cost else None), (self._info if want_info else None)

    def command_resident(self, sync=False):
        self._ck(self.fn["m3p2i_command_resident"](self.h, C.byref(self._info) if sync else None),
                 "m3p2i_command_resident")
        return self._info if sync else None

    def fetch_result(self, want_cost=True):
        self._ck(self.fn["m3p2i_fetch_result"](self.h, A.as_fp(self._act), A.as_fp(self._cost) if want_cost else None),
                 "m3p2i_fetch_result")
        return self._act, (self._cost if want_cost else None)

    def rollout_actions(self, actions):
        a = _f32(actions)
        if a.shape != (self.K, self.T, self.nu):
            raise ValueError(f"actions must be [{self.K},{self.T},{self.nu}]")
        st = np.empty((self.K, self.T, 4), np.float32)
        ch = np.empty((self.K, self.T), np.float32)
        self._ck(self.fn["m3p2i_rollout_actions"](self.h, A.as_fp(a), A.as_fp(st), A.as_fp(ch)), "m3p2i_rollout_actions")
        return st, ch

    def sample_actions(self):
        """Shift the stored sequences and return this tick's perturbed actions [K,T,nu] (generic callback path)."""
        out = np.empty((self.K, self.T, self.nu), np.float32)
        self._ck(self.fn["m3p2i_sample_actions"](self.h, A.as_fp(out)), "m3p2i_sample_actions")
        return out

    def update_only(self, cost_horizon, actions):
        ch, a = _f32(cost_horizon), _f32(actions)
        if ch.shape != (self.K, self.T) or a.shape != (self.K, self.T, self.nu):
            raise ValueError("cost_horizon must be [K,T] and actions [K,T,nu]")
        out = np.empty((self.T, self.nu), np.float32)
        info = A.CommandInfo()
        self._ck(self.fn["m3p2i_update_only"](self.h, A.as_fp(ch), A.as_fp(a), A.as_fp(out), C.byref(info)),
                 "m3p2i_update_only")
        return out, info

    def top_trajs(self, n=A.TOP_N):
        idx = np.empty(n, np.int32)
        w = np.empty(n, np.float32)
        tr = np.empty((n, self.T, 2), np.float32)
        self._ck(self.fn["m3p2i_top_trajs"](self.h, n, idx.ctypes.data_as(A.ip), A.as_fp(w), A.as_fp(tr)),
                 "m3p2i_top_trajs")
        return idx, w, tr

    _SHAPES = {A.BUF_ACTIONS: lambda s: (s.K, s.T, s.nu), A.BUF_STATES: lambda s: (s.K, s.T, 4),
               A.BUF_COST_HORIZON: lambda s: (s.K, s.T), A.BUF_COST_DISC: lambda s: (s.Kg,),
               A.BUF_COST_SUM: lambda s: (s.K,), A.BUF_WEIGHTS: lambda s: (3, s.Kg),
               A.BUF_NOISE: lambda s: (s.K, s.T, s.nu)}

    def read_buffer(self, which):
        out = np.empty(self._SHAPES[which](self), np.float32)
        self._ck(self.fn["m3p2i_read_buffer"](self.h, which, A.as_fp(out), out.size), "m3p2i_read_buffer")
        return out

    def get_buffer(self, which):
        ptr, nbytes = A.vp(), C.c_size_t()
        self._ck(self.fn["m3p2i_get_buffer"](self.h, which, C.byref(ptr), C.byref(nbytes)), "m3p2i_get_buffer")
        return ptr.value, nbytes.value

    # ---------------------------------------------------------------- K sharded over ranks
    def partials_len(self):
        return self.fn["m3p2i_partials_len"](self.h)

    def phase_rollout(self):
        J = np.empty(self.K, np.float32)
        self._ck(self.fn["m3p2i_phase_rollout"](self.h, A.as_fp(J)), "m3p2i_phase_rollout")
        return J

    def phase_partials(self, J_global):
        J = _f32(J_global)
        if J.size != self.Kg:
            raise ValueError(f"J_global needs {self.Kg} entries")
        out = np.empty(self.partials_len(), np.float32)
        self._ck(self.fn["m3p2i_phase_partials"](self.h, A.as_fp(J), A.as_fp(out)), "m3p2i_phase_partials")
        return out

    def phase_finish(self, partials_sum, want_cost=True):
        p = _f32(partials_sum)
        self._ck(self.fn["m3p2i_phase_finish"](self.h, A.as_fp(p), A.as_fp(self._act),
                                              A.as_fp(self._cost) if want_cost else None, C.byref(self._info)),
                 "m3p2i_phase_finish")
        return self._act, (self._cost if want_cost else None), self._info

    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._ck(self.fn["m3p2i_comm_init"](self.h, rank, nranks, C.cast(buf, A.vp)), "m3p2i_comm_init")

    def peer_export(self):
        """This rank's mailbox descriptor (bytes) for the exchange over NVLink peer memory (m3p2i_peer_export)."""
        ph = A.PeerHandle()
        self._ck(self.fn["m3p2i_peer_export"](self.h, C.byref(ph)), "m3p2i_peer_export")
        return bytes(ph)

    def peer_attach(self, rank, nranks, descriptors):
        """descriptors: the peer_export() bytes of every rank, indexed by rank. Follow with a host barrier."""
        if len(descriptors) != nranks or any(len(d) != C.sizeof(A.PeerHandle) for d in descriptors):
            raise ValueError(f"need {nranks} descriptors of {C.sizeof(A.PeerHandle)} bytes (peer_export() of every rank)")
        arr = (A.PeerHandle * nranks)(*[A.PeerHandle.from_buffer_copy(d) for d in descriptors])
        self._ck(self.fn["m3p2i_peer_attach"](self.h, rank, nranks, arr), "m3p2i_peer_attach")

    # ---------------------------------------------------------------- persistent K-env sim facade
    def sim_reset(self):
        self._ck(self.fn["m3p2i_sim_reset"](self.h), "m3p2i_sim_reset")

    def sim_set_velocity_target(self, u):
        a = _f32(u)
        if a.shape != (self.K, self.nu):
            raise ValueError(f"velocity targets must be [{self.K},{self.nu}]")
        self._ck(self.fn["m3p2i_sim_set_velocity_target"](self.h, A.as_fp(a)), "m3p2i_sim_set_velocity_target")

    def sim_apply_forces(self, f_robot, f_box):
        fr = None if f_robot is None else _f32(f_robot)
        fb = None if f_box is None else _f32(f_box)
        self._ck(self.fn["m3p2i_sim_apply_forces"](self.h, A.as_fp(fr), A.as_fp(fb)), "m3p2i_sim_apply_forces")

    def sim_step(self):
        self._ck(self.fn["m3p2i_sim_step"](self.h), "m3p2i_sim_step")

    def sim_write(self, dof, root):
        d = None if dof is None else _f32(dof)
        r = None if root is None else _f32(root)
        self._ck(self.fn["m3p2i_sim_write"](self.h, A.as_fp(d), A.as_fp(r)), "m3p2i_sim_write")

    def sim_cost(self):
        out = np.empty(self.K, np.float32)
        self._ck(self.fn["m3p2i_sim_cost"](self.h, A.as_fp(out)), "m3p2i_sim_cost")
        return out

    def sim_read(self):
        n_link = 1 if self.cfg.env_type == A.ENV_POINT else 3
        n_con = 1 if self.cfg.env_type == A.ENV_POINT else 3
        dof = np.empty((self.K, 2 * self.ndof), np.float32)
        root = np.empty((self.K, self.n_actors, 13), np.float32)
        link = np.empty((self.K, n_link, 13), np.float32)
        con = np.empty((self.K, n_con, 3), np.float32)
        self._ck(self.fn["m3p2i_sim_read"](self.h, A.as_fp(dof), A.as_fp(root), A.as_fp(link), A.as_fp(con)),
                 "m3p2i_sim_read")
        return dof, root, link, con


def comm_unique_id():
    L = lib()
    buf = (C.c_char * 128)()
    rc = L.fn["m3p2i_comm_unique_id"](C.cast(buf, A.vp))
    if rc != 0:
        raise NativeError("m3p2i_comm_unique_id failed: " + L.fn["m3p2i_last_error"]().decode(errors="replace"))
    return bytes(buf)


def device_count():
    return lib().fn["m3p2i_device_count"]()
