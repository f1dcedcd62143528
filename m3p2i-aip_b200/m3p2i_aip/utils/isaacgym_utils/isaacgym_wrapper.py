"""Sim facade with the accessor surface of the reference's IsaacGymWrapper, backed by the native B200 integrator.

Mirrors utils/isaacgym_utils/isaacgym_wrapper.py of the reference: IsaacGymConfig (:7-16), tensor views
`_dof_state [K, 2*ndof]`, `_root_state [K, n_actor, 13]` (:98-104), `robot_pos` / `robot_vel` (:120-126), the
name -> tensor getters (:128-188), the setters (:190-203), `update_dyn_obs` (:205-220) and `step()` (:354-360).
There is no IsaacGym / PhysX here: the K environments live in the native library (libm3p2i_b200.so) and the
tensors below are host mirrors refreshed on demand.

`backend` is any object with the planner/sim method surface of m3p2i_b200.native.NativePlanner; by default the
native CUDA backend is created (and its absence is an error: there is no CPU fallback in the product).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from m3p2i_b200 import _abi as A
from m3p2i_b200 import scene as S


@dataclass
class IsaacGymConfig():
    dt: float = 0.05
    substeps: int = 2
    use_gpu_pipeline: bool = True
    num_threads: int = 8
    viewer: bool = False
    spacing: float = 10
    camera_pos: List[float] = field(default_factory=lambda: [1.5, 6, 8])
    camera_target: List[float] = field(default_factory=lambda: [1.5, 0, 0])
    # not in the reference: after a fused command() the K envs hold the END states of the rollouts, as in the reference,
    # and the first access to a tensor view reads all of them back (2.5 MB at K=4096) -- which run_tamp then overwrites
    # with the real state (reactive_tamp.py:45-48). True skips that read-back: the views keep what was last written.
    skip_rollout_readback: bool = False


_LINKS = {"point_env": {("point_robot", "link_y"): 0},
          "panda_env": {("panda", "panda_leftfinger"): 0, ("panda", "panda_rightfinger"): 1, ("panda", "panda_hand"): 2}}
_CONTACTS = {"point_env": {"dyn-obs": 0}, "panda_env": {"table": 0, "shelf_stand": 1, "cubeB": 2}}


def _rows_equal(a):
    """True when every env row of a [K, ...] array equals row 0 (the `_dof_state[:] = one_state` broadcast of
    reactive_tamp.py:45-46): each row against its predecessor on the flat contiguous buffer -- 3x cheaper than comparing
    with a stride-0 broadcast view, and this check runs on every tick."""
    flat = np.ascontiguousarray(a).reshape(-1)
    row = flat.size // a.shape[0]
    return bool(np.array_equal(flat[row:], flat[:-row]))


class IsaacGymWrapper:
    def __init__(self, cfg: IsaacGymConfig, env_type: str = "point_env", num_envs: int = 1, viewer: bool = False,
                 device: str = "cuda:0", cube_on_shelf: bool = False, backend=None, backend_factory=None,
                 actors=None):
        self.cfg = cfg
        self.env_type = env_type
        self.num_envs = int(num_envs)
        self.device = "cpu"            # host mirrors; the device state is owned by the native library
        self.requested_device = device
        self.cube_on_shelf = cube_on_shelf
        self.viewer = None
        self.env_cfg = actors if actors is not None else S.default_actors(env_type)
        self.robot_indices = torch.tensor([i for i, a in enumerate(self.env_cfg) if a.type == "robot"])
        self.robot_per_env = len(self.robot_indices)
        self.dofs_per_robot = 2 if env_type == "point_env" else 9
        n_robot_bodies = 3 if env_type == "point_env" else 11
        self.bodies_per_env = len(self.env_cfg) - self.robot_per_env + n_robot_bodies * self.robot_per_env
        self.scene = S.build_point_scene(self.env_cfg) if env_type == "point_env" else S.build_panda_scene(self.env_cfg)
        self._backend_factory = backend_factory
        self.backend = backend
        K = self.num_envs
        self._dof0 = S.initial_dof_state(self.env_cfg)
        self._root0 = S.initial_root_state(self.env_cfg, cube_on_shelf)
        self.__dof = torch.from_numpy(np.tile(self._dof0, (K, 1)))
        self.__root = torch.from_numpy(np.tile(self._root0[None], (K, 1, 1)))
        n_link = 1 if env_type == "point_env" else 3
        self.__link = torch.zeros(K, n_link, 13)
        self.__contact = torch.zeros(K, len(_CONTACTS[env_type]), 3)
        self._host_dirty = False     # device state is newer than the host mirrors
        self._push_pending = True    # host mirrors (set by the caller) must be pushed before the next device op
        if self.backend is None and backend_factory is None:
            self._make_default_backend()

    # ------------------------------------------------------------------ backend management
    def _make_default_backend(self, planner_cfg=None):
        from m3p2i_b200 import native
        self.backend = native.NativePlanner.for_sim(self, planner_cfg)

    def attach_planner(self, cfg, noise_mode=A.NOISE_TABLE, seed=0):
        """Called by the planner (MPPI.__init__) when its callbacks are bound to this sim: re-creates the backend
        with the full planner configuration so that rollout and update run fused on the same K environments."""
        old = self.backend
        self._refresh()
        if self._backend_factory is not None:
            self.backend = self._backend_factory(self, cfg, noise_mode, seed)
        else:
            from m3p2i_b200 import native
            self.backend = native.NativePlanner.for_sim(self, cfg, noise_mode=noise_mode, seed=seed)
        if old is not None and old is not self.backend and hasattr(old, "close"):
            old.close()
        self._push_pending = True
        self._has_planner_cfg = True
        return self.backend

    def _ensure_backend(self):
        if self.backend is None:
            self.backend = self._backend_factory(self, None, A.NOISE_TABLE, 0)
        return self.backend

    def _push(self):
        """Host mirrors -> device envs (reactive_tamp.py:45-48)."""
        if not self._push_pending:
            return
        b = self._ensure_backend()
        dof, root = self.__dof.numpy(), self.__root.numpy()
        if self.num_envs == 1 or (_rows_equal(dof) and _rows_equal(root)):
            b.set_state(dof[0], root[0])
        else:
            b.set_state(dof[0], root[0])   # fixed actors / floating plate pose
            b.sim_write(dof, root)
        self._push_pending = False

    def _refresh(self):
        """Device envs -> host mirrors (the refresh_*_tensor calls of isaacgym_wrapper.py:357-360)."""
        if not self._host_dirty or self.backend is None:
            return
        dof, root, link, con = self.backend.sim_read()
        self.__dof.copy_(torch.from_numpy(dof))
        self.__root.copy_(torch.from_numpy(root))
        self.__link.copy_(torch.from_numpy(link))
        self.__contact.copy_(torch.from_numpy(con))
        self._host_dirty = False

    def mark_device_advanced(self):
        """The fused command() moved the K envs on the device."""
        if getattr(self.cfg, "skip_rollout_readback", False):
            self._push_pending = True   # the device envs no longer match the host mirrors: re-push before the next device op
        else:
            self._host_dirty = True

    # ------------------------------------------------------------------ tensor views
    @property
    def _dof_state(self):
        self._refresh()
        return self.__dof

    @property
    def _root_state(self):
        self._refresh()
        return self.__root

    @property
    def _rigid_body_state(self):
        self._refresh()
        return self.__link

    @property
    def _net_contact_force(self):
        self._refresh()
        return self.__contact

    @property
    def robot_pos(self):
        return torch.index_select(self._dof_state, 1, torch.tensor([0, 2]))

    @property
    def robot_vel(self):
        return torch.index_select(self._dof_state, 1, torch.tensor([1, 3]))

    def _get_actor_index_by_name(self, name: str):
        return torch.tensor([a.name for a in self.env_cfg].index(name))

    def _get_actor_index_by_robot_index(self, robot_idx: int):
        return self.robot_indices[robot_idx]

    def get_actor_position_by_actor_index(self, actor_idx):
        return self._root_state[:, int(actor_idx), 0:3]

    def get_actor_position_by_name(self, name: str):
        return self.get_actor_position_by_actor_index(self._get_actor_index_by_name(name))

    def get_actor_position_by_robot_index(self, robot_idx: int):
        return self.get_actor_position_by_actor_index(self._get_actor_index_by_robot_index(robot_idx))

    def get_actor_velocity_by_actor_index(self, idx):
        return self._root_state[:, int(idx), 7:10]

    def get_actor_velocity_by_name(self, name: str):
        return self.get_actor_velocity_by_actor_index(self._get_actor_index_by_name(name))

    def get_actor_velocity_by_robot_index(self, robot_idx: int):
        return self.get_actor_velocity_by_actor_index(self._get_actor_index_by_robot_index(robot_idx))

    def get_actor_orientation_by_actor_index(self, idx):
        return self._root_state[:, int(idx), 3:7]

    def get_actor_orientation_by_name(self, name: str):
        return self.get_actor_orientation_by_actor_index(self._get_actor_index_by_name(name))

    def get_actor_orientation_by_robot_index(self, robot_idx: int):
        return self.get_actor_orientation_by_actor_index(self._get_actor_index_by_robot_index(robot_idx))

    def get_actor_link_by_name(self, actor_name: str, link_name: str):
        """[K,13] rigid-body rows; single-body actors (link 'box') are their root rows."""
        key = (actor_name, link_name)
        links = _LINKS[self.env_type]
        if key in links:
            return self._rigid_body_state[:, links[key], :]
        if link_name == "box":
            return self._root_state[:, int(self._get_actor_index_by_name(actor_name)), :]
        raise KeyError(f"link {link_name!r} of actor {actor_name!r} is not modelled by the native integrator")

    def get_actor_contact_forces_by_name(self, actor_name: str, link_name: str):
        contacts = _CONTACTS[self.env_type]
        if actor_name not in contacts:
            raise KeyError(f"net contact force of {actor_name!r} is not tracked by the native integrator")
        return self._net_contact_force[:, contacts[actor_name]]

    # ------------------------------------------------------------------ setters
    def set_dof_state_tensor(self, u):
        # bring BOTH host mirrors up to date first: the next push writes dof and root together, and Isaac Gym's
        # set_dof_state_tensor never touches the root states (a stale root mirror would rewind every actor)
        self._refresh()
        u = torch.as_tensor(u, dtype=torch.float32).reshape(self.num_envs, -1)
        if u.data_ptr() != self.__dof.data_ptr():
            self.__dof.copy_(u)
        self._push_pending = True
        self._host_dirty = False

    def set_actor_root_state_tensor(self, u):
        self._refresh()   # see set_dof_state_tensor: the dof mirror must not be stale when the pair is pushed
        u = torch.as_tensor(u, dtype=torch.float32).reshape(self.num_envs, -1, 13)
        if u.data_ptr() != self.__root.data_ptr():
            self.__root.copy_(u)
        self._push_pending = True
        self._host_dirty = False

    def set_dof_velocity_target_tensor(self, u):
        self._push()
        u = torch.as_tensor(u, dtype=torch.float32).reshape(-1, self.dofs_per_robot)
        if u.shape[0] == 1 and self.num_envs > 1:
            u = u.expand(self.num_envs, -1)
        self.backend.sim_set_velocity_target(u.contiguous().numpy())

    def set_dof_actuation_force_tensor(self, u):
        raise NotImplementedError("effort drive mode is not used by any shipped configuration")

    def apply_rigid_body_force_tensors(self, u):
        """forces [K, bodies_per_env, 3]: the suction pair written by calculate_suction (skill_utils.py:84-90):
        row `box` actor index and the last body row (robot)."""
        if self.env_type != "point_env":
            raise NotImplementedError("external body forces are only modelled in point_env (suction)")
        self._push()
        f = torch.as_tensor(u, dtype=torch.float32).reshape(self.num_envs, -1, 3)
        bi = int(self._get_actor_index_by_name("box"))
        self.backend.sim_apply_forces(f[:, -1, :2].contiguous().numpy(), f[:, bi, :2].contiguous().numpy())

    def update_dyn_obs(self, i, period=100):
        """isaacgym_wrapper.py:205-220: teleports the dynamic obstacle on a triangle wave."""
        idx = int(self._get_actor_index_by_name("dyn-obs"))
        root = self._root_state
        off = torch.tensor([0.01, 0.01, 0.0]) if self.env_type == "point_env" else torch.zeros(3)
        if i % period > period / 4 and i % period < period / 4 * 3:
            root[:, idx, :3] += off
        else:
            root[:, idx, :3] -= off
        self.set_actor_root_state_tensor(root)

    def play_with_cube(self):
        pass  # keyboard interaction of the reference viewer (isaacgym_wrapper.py:421-433): no viewer here

    def visualize_trajs(self, trajs):
        """The reference draws the planner's top trajectories as lines in its viewer (isaacgym_wrapper.py:374-391,
        sim.py:54-56). There is no viewer here: the trajectories [n, T, 2 or 3] are kept for the caller to plot."""
        self.last_trajs = trajs.detach().cpu().clone() if torch.is_tensor(trajs) else np.array(trajs, copy=True)

    def initialize_keyboard_listeners(self):
        pass  # viewer-only (isaacgym_wrapper.py:393-405)

    def keyboard_control(self):
        pass  # viewer-only (isaacgym_wrapper.py:407-419)

    def step(self):
        self._push()
        self.backend.sim_step()
        self._host_dirty = True

    def stop_sim(self):
        if self.backend is not None and hasattr(self.backend, "close"):
            self.backend.close()
        self.backend = None
