"""MPPI controller with the reference's API, running on the native B200 hot path.

Mirrors planners/motion_planner/mppi.py: MPPIConfig (:9-59), MPPI.__init__ (:82-203), command (:211-264) and the
attributes a caller reads afterwards (weights, cost_total, states, actions, mean_action, best_traj*, top_trajs,
delta). Two execution paths:

* fused (default when `dynamics` / `running_cost` are bound methods of an object whose `.sim` is this package's
  IsaacGymWrapper and `.objective` this package's Objective -- exactly what scripts/reactive_tamp.py:37-41,63-73
  passes): one m3p2i_command() call = noise, perturbation, T-step rollout, costs, softmin, mean update on the GPU.
  The Python callbacks are not invoked.
* generic (any other callables): the reference's T-step Python loop (mppi.py:296-315) calls the user's dynamics
  and running_cost; perturbation and the softmin / weighted update (mppi.py:381-416, 430-456, 485-518) go through
  m3p2i_update_only().
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch

from m3p2i_b200 import _abi as A
from m3p2i_b200 import native
from m3p2i_b200 import scene as S
from m3p2i_aip.utils import mppi_utils


@dataclass
class MPPIConfig(object):
    num_samples: int = 200
    horizon: int = 12
    nx: int = 4
    mppi_mode: str = 'halton-spline'
    sampling_method: str = "halton"   # "halton" | "random" | in-kernel counter RNG: "philox" (white), "philox-spline" (smooth)
    noise_sigma: Optional[List[List[float]]] = None
    noise_mu: Optional[List[float]] = None
    device: str = "cuda:0"
    lambda_: float = 1.0
    update_lambda: bool = False
    update_cov: bool = False
    u_min: Optional[List[float]] = None
    u_max: Optional[List[float]] = None
    u_init: float = 0.0
    U_init: Optional[List[List[float]]] = None
    u_scale: float = 1
    u_per_command: int = 1
    rollout_var_discount: float = 0.95
    sample_null_action: bool = False
    sample_previous_plan: bool = True
    sample_other_priors: bool = False
    noise_abs_cost: bool = False
    filter_u: bool = False
    use_priors: bool = False
    seed_val: int = 0
    eta_u_bound: int = 10
    eta_l_bound: int = 5


_SEQ_KEYS = ("mean_action", "mean_action_1", "mean_action_2", "best_traj", "best_traj_1", "best_traj_2")


class MPPI():
    def __init__(self, cfg, dynamics: Callable, running_cost: Callable):
        self.env_type = cfg.env_type
        self.multi_modal = bool(cfg.multi_modal)
        self._full_cfg = cfg
        m = cfg.mppi
        self.mppi_mode = getattr(m, "mppi_mode", "halton-spline")
        if self.mppi_mode not in ("halton-spline", "simple"):
            raise ValueError(f"unknown mppi_mode {self.mppi_mode!r}")
        self.sampling_method = getattr(m, "sampling_method", "halton")
        if self.sampling_method not in ("halton", "random", "philox", "philox-spline"):
            raise ValueError(f"unknown sampling_method {self.sampling_method!r}")
        # covariance adaptation (mppi.py:43,508-516) runs in the update kernel; like the reference it only exists on the
        # single-mode path (M3P2I._update_multi_modal_distribution has no covariance branch)
        self.update_cov = bool(getattr(m, "update_cov", False))
        self.K = int(m.num_samples)
        self.half_K = int(self.K / 2)
        self.T = int(m.horizon)
        self.filter_u = bool(getattr(m, "filter_u", False))
        self.lambda_ = getattr(m, "lambda_", 1.0)
        self.sample_null_action = bool(getattr(m, "sample_null_action", False))
        self.u_per_command = getattr(m, "u_per_command", 1)
        self.device = "cpu"
        self.u_scale = float(getattr(m, "u_scale", 1))
        if not getattr(m, "noise_sigma", None):
            m.noise_sigma = np.identity(int(m.nx / 2)).tolist()
        self.noise_sigma = torch.tensor(m.noise_sigma, dtype=torch.float32)
        self.nx = m.nx
        self.nu = self.noise_sigma.shape[0]
        if self.K < 20:
            raise ValueError("num_samples must be >= 20 (command() takes the top 20 trajectories, mppi.py:248)")
        if self.filter_u and self.T < 9:
            raise ValueError("filter_u needs horizon >= 9 (Savitzky-Golay window, mppi.py:190,259)")
        self.F = dynamics
        self.running_cost = running_cost
        self.step_size_mean = 0.98
        self.gamma = getattr(m, "rollout_var_discount", 0.95)
        self.sgf_window, self.sgf_order = 9, 2
        self.seed_val = int(getattr(m, "seed_val", 0))
        self._delta = None
        self._delta_uploaded = False
        self._delta_on_device = False
        self.knot_scale, self.degree = 4, 2           # mppi.py:169,173
        # digit permutations of the generalised Halton sequence, uint16 [n_knots * nu, stride] (ghalton.EA_PERMS when the
        # package is installed, mppi_utils.py:88-95); None = the plain sequence (the reference's use_ghalton=False branch)
        self.halton_perms = getattr(m, "halton_perms", None)
        self._info = None
        self.gripper_command = None
        self.state = None

        owner = getattr(dynamics, "__self__", None)
        sim = getattr(owner, "sim", None)
        obj = getattr(owner, "objective", None)
        from m3p2i_aip.planners.motion_planner.cost_functions import Objective
        from m3p2i_aip.utils.isaacgym_utils.isaacgym_wrapper import IsaacGymWrapper
        self.noise_abs_cost = bool(getattr(m, "noise_abs_cost", False))
        self.noise_sigma_inv = torch.inverse(self.noise_sigma)
        mu = getattr(m, "noise_mu", None)
        self.noise_mu = torch.tensor(mu, dtype=torch.float32) if mu else torch.zeros(self.nu)
        self._rng = torch.Generator().manual_seed(self.seed_val)
        self.U = None
        # classic MPPI resamples its noise every call and runs through the callbacks (the caller's cfg is not touched)
        fused_opt = self.mppi_mode != "simple" and bool(getattr(m, "fused", True))
        self.fused = (isinstance(sim, IsaacGymWrapper) and isinstance(obj, Objective)
                      and getattr(running_cost, "__self__", None) is owner and fused_opt
                      and sim.num_envs == self.K)
        noise_mode = {"philox": A.NOISE_PHILOX, "philox-spline": A.NOISE_PHILOX_SPLINE}.get(self.sampling_method, A.NOISE_TABLE)
        if self.fused:
            self._sim, self._objective = sim, obj
            self.backend = sim.attach_planner(cfg, noise_mode=noise_mode, seed=self.seed_val)
        else:
            self._sim, self._objective = None, None
            if isinstance(sim, IsaacGymWrapper) and sim.num_envs == self.K:
                # the callbacks will step this sim themselves; give its cost / suction kernels the full configuration
                sim.attach_planner(cfg, noise_mode=A.NOISE_TABLE, seed=self.seed_val)
            scene = S.build_point_scene() if self.env_type == "point_env" else S.build_panda_scene()
            self.backend = native.NativePlanner(S.build_config(cfg, noise_mode=noise_mode, seed=self.seed_val), scene)
            if self.filter_u:
                self.backend.set_filter_matrix(S.savgol_matrix(self.T))
            # the softmin/update kernels need no scene state, but the handle wants one
            actors = S.default_actors(self.env_type)
            self.backend.set_state(S.initial_dof_state(actors), S.initial_root_state(actors))
            if self.mppi_mode == "simple":
                # classic MPPI update = softmin of the TOTAL cost at temperature lambda and U <- sum_k w_k a_k
                # (U + sum w (a - U) with sum w = 1): the same kernels with gamma = 1, step size 1, beta = lambda
                c = S.build_config(cfg, noise_mode=A.NOISE_TABLE, seed=self.seed_val)
                c.gamma, c.step_size_mean, c.multi_modal, c.filter_u = 1.0, 1.0, 0, 0
                self.backend.close()
                self.backend = native.NativePlanner(c, scene)
                self.backend.set_state(S.initial_dof_state(actors), S.initial_root_state(actors))
                self.U = self._sample_noise(self.T)

    # ------------------------------------------------------------------ noise table (mppi.py:386-392,458-483)
    @property
    def delta(self):
        if self._delta is None and self._delta_on_device:
            return torch.from_numpy(self.backend.get_noise())   # the table lives on the device; read back on demand
        return self._delta

    @delta.setter
    def delta(self, value):
        self._delta = value
        self._delta_uploaded = False
        self._delta_on_device = False

    def get_samples(self, sample_shape, **kwargs):
        if self.sampling_method == "halton":
            return torch.from_numpy(mppi_utils.halton_spline_table(sample_shape, self.T, self.nu))
        if self.sampling_method == "random":
            # noise_dist.sample((K, T)) of mppi.py:479-480: fresh N(noise_mu, noise_sigma) draws on every command,
            # from the planner's own generator (full covariance through its Cholesky factor)
            return self._sample_noise(sample_shape, self.T) + self.noise_mu
        raise ValueError(f"unknown sampling_method {self.sampling_method!r}")

    def _sample_noise(self, *shape):
        """N(0, noise_sigma) draws [*shape, nu] (MultivariateNormal of mppi.py:129-131, own generator)."""
        L = torch.linalg.cholesky(self.noise_sigma)
        return torch.randn(*shape, self.nu, generator=self._rng) @ L.T

    def _ensure_noise(self):
        if self.sampling_method in ("philox", "philox-spline") and self._delta is None:
            return
        if self.sampling_method == "halton" and self._delta is None:
            # the once-sampled table of mppi.py:458-478 (Halton knots, erfinv, one smoothing spline per sample and
            # dimension) is built by the backend for its own shard: on the device, no K * nu scipy calls
            if not self._delta_on_device:
                self.backend.set_noise_halton_spline(self.knot_scale, self.degree, 0.5, perms=self.halton_perms)
                self._delta_on_device = True
            return
        if self.sampling_method == "random" or self._delta is None:
            self.delta = self.get_samples(self.K, base_seed=0)
        if not self._delta_uploaded:
            d = self._delta.detach().cpu().numpy() if torch.is_tensor(self._delta) else np.asarray(self._delta)
            if self.backend.cfg.noise_mode != A.NOISE_TABLE:
                raise RuntimeError("a delta table was given but the planner was created with an in-kernel sampling_method")
            self.backend.set_noise_table(d)
            self._delta_uploaded = True

    # ------------------------------------------------------------------ command (mppi.py:211-264)
    def command(self, state):
        if not torch.is_tensor(state):
            state = torch.tensor(state)
        self.state = state.to(torch.float32)
        self._lazy = {}
        if self.mppi_mode == "simple":
            action, info = self._command_simple()
            self._info = info
            return action
        self._ensure_noise()
        if self.fused:
            obj = self._objective
            if obj.task is None:
                raise RuntimeError("Objective.update_objective(task, goal) must be called before command()")
            self._sim._push()
            grip = self.gripper_command if self.env_type == "panda_env" else None
            self.backend.set_objective(obj.task, obj.goal_array(), grip)
            action, _, info = self.backend.command(want_cost=False)
            self._sim.mark_device_advanced()
        else:
            action, info = self._command_generic()
        self._info = info
        return torch.from_numpy(np.array(action, copy=True))

    def _dynamics(self, state, u, t=None):
        return self.F(state, u, t=None)

    def _running_cost(self, state):
        return self.running_cost(state)

    def _rollout_callbacks(self, act):
        """The reference's T-step loop over the user's callbacks (mppi.py:296-315): -> cost_horizon, states, actions"""
        K, T = self.K, self.T
        state = self.state.view(1, -1).repeat(K, 1) if self.state.shape != (K, self.nx) else self.state
        cost_horizon = torch.zeros(K, T)
        states, actions = [], []
        for t in range(T):
            u = self.u_scale * act[:, t]
            if self.sample_null_action:
                u[K - 1] = 0.0
            state, u = self._dynamics(state, u, t)
            c = self._running_cost(state)
            cost_horizon[:, t] = torch.as_tensor(c, dtype=torch.float32)
            states.append(torch.as_tensor(state, dtype=torch.float32))
            actions.append(torch.as_tensor(u, dtype=torch.float32))
        return cost_horizon, torch.stack(states, dim=-2), torch.stack(actions, dim=-2)

    def _command_simple(self):
        """mppi_mode='simple' (mppi.py:220-233,335-363): fresh Gaussian noise, lambda * U Sigma^-1 eps control cost,
        softmin of the total cost at temperature lambda. Rollouts run through the callbacks, the softmin and the
        weighted sum through m3p2i_update_only."""
        K, T, nu = self.K, self.T, self.nu
        self.U = torch.roll(self.U, -1, dims=0)
        noise = self._sample_noise(K, T)
        u_min = torch.tensor(list(self.backend.cfg.u_min)[:nu])
        u_max = torch.tensor(list(self.backend.cfg.u_max)[:nu])
        act = torch.max(torch.min(self.U + noise, u_max), u_min)
        if self.env_type == "panda_env" and self.gripper_command in ("open", "close"):
            act[:, :, 7:9] = 1.5 if self.gripper_command == "open" else -1.5
        cost_horizon, states, actions = self._rollout_callbacks(act)
        if self.sample_null_action:
            act[K - 1] = 0.0                              # perturbed_action[K-1] is overwritten by the null action (mppi.py:302)
        cs = cost_horizon.sum(1)
        cost_total = cs + cs.mean()                      # aliasing quirk, mppi.py:282-284,325
        noise = act - self.U                              # bounded noise, mppi.py:356
        ac = self.lambda_ * (noise.abs() if self.noise_abs_cost else noise) @ self.noise_sigma_inv
        cost_total = cost_total + torch.sum(self.U * ac, dim=(1, 2))
        st = self.backend.get_planner_state()
        st.beta = float(self.lambda_)
        self.backend.set_planner_state(st)
        packed = torch.zeros(K, T)
        packed[:, 0] = cost_total
        new_U, info = self.backend.update_only(packed.numpy(), act.numpy())
        self.U = torch.from_numpy(np.array(new_U, copy=True))
        self._lazy.update(states=states, actions=actions / self.u_scale, cost_total=cost_total)
        action = self.U[: self.u_per_command].clone()
        if self.filter_u and action.shape[0] >= 9:
            action = torch.from_numpy(S.savgol_matrix(action.shape[0]) @ action.numpy())
        return action, info

    def _command_generic(self):
        """The reference's control flow with the user's callbacks (mppi.py:237-245,381-416,275-332): shift and
        perturbation (m3p2i_sample_actions), T callback steps, softmin / mean update / filter (m3p2i_update_only)."""
        grip = self.gripper_command if self.env_type == "panda_env" else None
        self.backend.set_objective("reach" if self.env_type == "panda_env" else "navigation",
                                   np.zeros(7 if self.env_type == "panda_env" else 2, np.float32), grip)
        act = torch.from_numpy(self.backend.sample_actions())
        cost_horizon, states, actions = self._rollout_callbacks(act)
        self._lazy["states"] = states
        _, info = self.backend.update_only(cost_horizon.numpy(), actions.numpy())
        cs = cost_horizon.sum(1)
        self._lazy["cost_total"] = cs + cs.mean()
        self._lazy["actions"] = actions / self.u_scale
        action = np.array(self.backend.fetch_result(want_cost=False)[0], copy=True)   # filtered when filter_u
        return action, info

    # ------------------------------------------------------------------ results of the last command (read lazily)
    def _seq(self, key):
        st = self.backend.get_planner_state()
        return torch.from_numpy(np.asarray(getattr(st, key)[: self.T * self.nu], np.float32).reshape(self.T, self.nu).copy())

    def _buf(self, name, which):
        if name not in self._lazy:
            self._lazy[name] = torch.from_numpy(self.backend.read_buffer(which))
        return self._lazy[name]

    mean_action = property(lambda s: s._seq("mean_action"))
    mean_action_1 = property(lambda s: s._seq("mean_action_1"))
    mean_action_2 = property(lambda s: s._seq("mean_action_2"))
    best_traj = property(lambda s: s._seq("best_traj"))
    best_traj_1 = property(lambda s: s._seq("best_traj_1"))
    best_traj_2 = property(lambda s: s._seq("best_traj_2"))

    @property
    def beta(self):
        return self.backend.get_planner_state().beta

    @property
    def cov_action(self):
        """per-dimension noise variance (mppi.py:175,514-515); constant unless update_cov"""
        st = self.backend.get_planner_state()
        return torch.tensor([st.cov_action[d] for d in range(self.nu)], dtype=torch.float32)

    @property
    def scale_tril(self):
        return torch.sqrt(self.cov_action)

    @property
    def weights(self):
        return self._buf("weights3", A.BUF_WEIGHTS)[0]

    @property
    def weights_1(self):
        return self._buf("weights3", A.BUF_WEIGHTS)[1, :self.half_K]

    @property
    def weights_2(self):
        return self._buf("weights3", A.BUF_WEIGHTS)[2, self.half_K:]

    @property
    def states(self):
        return self._buf("states", A.BUF_STATES)

    @property
    def actions(self):
        return self._buf("actions", A.BUF_ACTIONS)

    @property
    def cost_total(self):
        if "cost_total" not in self._lazy:
            self._lazy["cost_total"] = torch.from_numpy(np.array(self.backend.fetch_result(want_cost=True)[1], copy=True))
        return self._lazy["cost_total"]

    def _topk(self):
        if "top" not in self._lazy:
            idx, w, _ = self.backend.top_trajs(20)
            if self.fused:
                _, _, tr = self.backend.top_trajs(20)
            else:
                tr = self._lazy["states"][torch.from_numpy(idx.astype(np.int64))][:, :, [0, 2]].numpy()
            self._lazy["top"] = (torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(w), torch.from_numpy(np.array(tr)))
        return self._lazy["top"]

    top_idx = property(lambda s: s._topk()[0])
    top_values = property(lambda s: s._topk()[1])
    top_trajs = property(lambda s: s._topk()[2])
