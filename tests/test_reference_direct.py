"""Randomised, direct comparisons with the UNMODIFIED reference code (tiers T1 and T2), beyond the fixed golden
cases: only in the build container (needs /root/reference; skipped elsewhere).

T2  cost parity : the reference's Objective.compute_cost on randomly posed K-env sims (this repo's facade on the oracle
                  integrator: random joint states, random cube poses incl. orientation, one random step so that contact
                  forces and velocities are populated) against the oracle's per-step cost for every task.
T1  update parity: random cost_horizon / actions through the reference's _update_distribution /
                  _update_multi_modal_distribution against the oracle's update (softmin, beta search, means, best rows).
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle_py as O
from helpers import assert_close, make_backend
from m3p2i_b200 import _abi as A
from m3p2i_b200 import scene as S

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    ref_m3p2i, ref_cost = G.import_reference()
    return G, ref_m3p2i, ref_cost


def _random_sim(env, cfg, K, rng, shelf=False):
    from m3p2i_aip.utils.isaacgym_utils import isaacgym_wrapper as wrapper
    sim = wrapper.IsaacGymWrapper(cfg.isaacgym, env, num_envs=K, device="cpu", cube_on_shelf=shelf,
                                  backend_factory=O.Oracle.for_sim)
    sim.attach_planner(cfg)
    actors = S.default_actors(env)
    dof = np.tile(S.initial_dof_state(actors), (K, 1)).astype(np.float32)
    root = np.tile(S.initial_root_state(actors, shelf)[None], (K, 1, 1)).astype(np.float32)
    if env == "point_env":
        dof[:, 0::2] = rng.uniform(-2.5, 2.5, (K, 2))
        dof[:, 1::2] = rng.uniform(-1.0, 1.0, (K, 2))
        for name in ("box", "dyn-obs"):
            i = S.actor_index(actors, name)
            root[:, i, :2] = rng.uniform(-2.5, 2.5, (K, 2))
            th = rng.uniform(-np.pi, np.pi, K)
            root[:, i, 3:7] = np.stack([0 * th, 0 * th, np.sin(th / 2), np.cos(th / 2)], 1)
        # half of the robots next to the block so that suction / pushing contacts occur
        near = rng.random(K) < 0.5
        bi = S.actor_index(actors, "box")
        ang = rng.uniform(-np.pi, np.pi, K)
        rad = rng.uniform(0.42, 0.7, K)
        dof[near, 0] = (root[:, bi, 0] + rad * np.cos(ang))[near]
        dof[near, 2] = (root[:, bi, 1] + rad * np.sin(ang))[near]
    else:
        lo = np.array([-2.8, -1.7, -2.8, -3.0, -2.8, 0.0, -2.8, 0.0, 0.0], np.float32)
        hi = np.array([2.8, 1.7, 2.8, -0.1, 2.8, 3.7, 2.8, 0.04, 0.04], np.float32)
        dof[:, 0::2] = rng.uniform(lo, hi, (K, 9))
        for name in ("cubeA", "cubeB"):
            i = S.actor_index(actors, name)
            root[:, i, :2] += rng.uniform(-0.15, 0.15, (K, 2))
            root[:, i, 2] += rng.uniform(0.0, 0.2, K)
            q = rng.standard_normal((K, 4))
            root[:, i, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    sim._dof_state[:] = torch.from_numpy(dof)
    sim._root_state[:] = torch.from_numpy(root)
    sim.set_dof_state_tensor(sim._dof_state)
    sim.set_actor_root_state_tensor(sim._root_state)
    nu = 2 if env == "point_env" else 9
    sim.set_dof_velocity_target_tensor(torch.from_numpy(rng.uniform(-1.5, 1.5, (K, nu)).astype(np.float32)))
    sim.step()
    return sim


POINT_GOAL = [-1.2, 0.8]
CASES = [("point_env", "navigation", False, False), ("point_env", "push", False, False), ("point_env", "pull", False, False),
         ("point_env", "push_pull", True, False), ("panda_env", "reach", False, False), ("panda_env", "reach", True, True),
         ("panda_env", "pick", False, False), ("panda_env", "place", False, False)]


@pytest.mark.parametrize("env,task,mm,shelf", CASES)
def test_cost_matches_reference_on_random_states(ref, env, task, mm, shelf):
    G, ref_m3p2i, ref_cost = ref
    from m3p2i_aip.planners.motion_planner.cost_functions import Objective
    K = 96
    rng = np.random.default_rng(100 + CASES.index((env, task, mm, shelf)))
    goal = POINT_GOAL if env == "point_env" else list(rng.uniform(-0.3, 0.3, 3) + np.array([0.2, 0.2, 1.1])) + [0.0, 0.0, 0.0, 1.0]
    cfg = S.make_cfg(env, task, goal, K, 12, multi_modal=mm, cube_on_shelf=shelf, device="cpu")
    sim = _random_sim(env, cfg, K, rng, shelf)
    g = torch.tensor(goal, dtype=torch.float32)
    theirs = ref_cost.Objective(cfg)
    theirs.update_objective(task, g)
    ours = Objective(cfg)
    ours.update_objective(task, g)
    c_ours = ours.compute_cost(sim).numpy().copy()
    c_ref = theirs.compute_cost(sim).numpy().copy()
    assert np.isfinite(c_ref).all()
    # the 1000-cost is a threshold on a contact force (cost_functions.py:165-169): a force within rounding of 0.1 N
    # may fall on either side; everything else agrees to fp32 rounding
    assert_close(c_ours, c_ref, 2e-4, 2e-4, f"{env} {task} mm={mm} cost", 0.011)
    assert np.ptp(c_ref) > 1e-3            # the random states actually exercise the cost


@pytest.mark.parametrize("env,mm", [("point_env", False), ("panda_env", False), ("point_env", True), ("panda_env", True)])
def test_update_matches_reference_on_random_costs(ref, env, mm):
    G, ref_m3p2i, ref_cost = ref
    K, T = 128, 12
    task = {"point_env": "push_pull" if mm else "push", "panda_env": "reach"}[env]
    cfg = S.make_cfg(env, task, POINT_GOAL if env == "point_env" else None, K, T, multi_modal=mm, device="cpu")
    cfg.mppi.filter_u = False
    torch.manual_seed(0)
    planner = ref_m3p2i.M3P2I(cfg, dynamics=lambda s, u, t=None: (s, u), running_cost=lambda s: torch.zeros(K))
    nu = planner.nu
    rng = np.random.default_rng(3 + 2 * mm + (env == "panda_env"))
    o = make_backend(O.Oracle, cfg)
    actors = S.default_actors(env)
    o.set_state(S.initial_dof_state(actors), S.initial_root_state(actors))
    for rep in range(3):
        scale = [0.5, 5.0, 50.0][rep]
        cost_h = (rng.random((K, T)) * scale + rng.random((K, 1)) * scale).astype(np.float32)
        actions = rng.uniform(-1.0, 1.0, (K, T, nu)).astype(np.float32)
        mean_before = planner.mean_action.clone()
        if mm:
            planner._update_multi_modal_distribution(torch.from_numpy(cost_h), torch.from_numpy(actions))
        else:
            planner._update_distribution(torch.from_numpy(cost_h), torch.from_numpy(actions))
        st = o.get_planner_state()
        np.asarray(st.mean_action)[: T * nu] = mean_before.numpy().ravel()
        o.set_planner_state(st)
        mean, info = o.update_only(cost_h, actions)
        w = o.read_buffer(A.BUF_WEIGHTS)
        assert_close(w[0], planner.weights.numpy(), 2e-3, 1e-7, f"{env} mm={mm} weights [{rep}]")
        assert_close(mean, planner.mean_action.numpy(), 2e-4, 2e-5, f"{env} mm={mm} mean_action [{rep}]")
        st = o.get_planner_state()
        if mm:
            half = K // 2
            assert_close(w[1, :half], planner.weights_1.numpy(), 2e-3, 1e-7, "weights_1")
            assert_close(w[2, half:], planner.weights_2.numpy(), 2e-3, 1e-7, "weights_2")
            for key in ("mean_action_1", "mean_action_2", "best_traj_1", "best_traj_2"):
                got = np.asarray(getattr(st, key)[: T * nu], np.float32).reshape(T, nu)
                assert_close(got, getattr(planner, key).numpy(), 2e-4, 2e-5, key)
            assert int(info.weight_pull > info.weight_push) == int(planner.get_pull_preference())
        else:
            got = np.asarray(st.best_traj[: T * nu], np.float32).reshape(T, nu)
            assert_close(got, planner.best_traj.numpy(), 0, 0, "best_traj")
            if env == "panda_env":
                assert st.beta == pytest.approx(float(planner.beta), rel=1e-12)
    o.close()


@pytest.mark.parametrize("env", ["point_env", "panda_env"])
def test_update_cov_matches_reference(ref, env):
    """mppi.update_cov=True (mppi.py:508-516): the per-dimension variance update from the weighted second moment of the
    samples, step 0.7, kappa 0.005, and the noise scale sqrt(cov) the NEXT command perturbs with (mppi.py:394,516):
    three successive random updates through the reference's _update_distribution against the oracle's update."""
    G, ref_m3p2i, ref_cost = ref
    K, T = 128, 12
    cfg = S.make_cfg(env, "push" if env == "point_env" else "reach", POINT_GOAL if env == "point_env" else None, K, T, device="cpu")
    cfg.mppi.filter_u = False
    cfg.mppi.update_cov = True
    torch.manual_seed(0)
    planner = ref_m3p2i.M3P2I(cfg, dynamics=lambda s, u, t=None: (s, u), running_cost=lambda s: torch.zeros(K))
    nu = planner.nu
    rng = np.random.default_rng(17)
    o = make_backend(O.Oracle, cfg)
    actors = S.default_actors(env)
    o.set_state(S.initial_dof_state(actors), S.initial_root_state(actors))
    st = o.get_planner_state()
    assert_close(np.asarray(st.cov_action[:nu]), planner.cov_action.numpy(), 1e-6, 0, "initial cov_action")
    for rep in range(3):
        cost_h = (rng.random((K, T)) * 5 + rng.random((K, 1)) * 5).astype(np.float32)
        actions = rng.uniform(-2.0, 2.0, (K, T, nu)).astype(np.float32)
        planner._update_distribution(torch.from_numpy(cost_h), torch.from_numpy(actions))
        mean, info = o.update_only(cost_h, actions)
        st = o.get_planner_state()
        assert_close(mean, planner.mean_action.numpy(), 2e-4, 2e-5, f"{env} mean_action [{rep}]")
        assert_close(np.asarray(st.cov_action[:nu]), planner.cov_action.numpy(), 2e-4, 1e-6, f"{env} cov_action [{rep}]")
        assert_close(np.sqrt(np.asarray(st.cov_action[:nu])), planner.scale_tril.numpy(), 2e-4, 1e-6, f"{env} scale_tril [{rep}]")
    # the adapted scale is what the next perturbation uses (mppi.py:394): sampled action = clamp(shifted mean + delta * sqrt(cov))
    delta = rng.standard_normal((K, T, nu)).astype(np.float32) * 0.1
    o.set_noise_table(delta)
    o.set_objective("push" if env == "point_env" else "reach", np.asarray(POINT_GOAL if env == "point_env" else [0.0] * 7, np.float32), None)
    mean = planner.mean_action.numpy()
    o.command()
    act = o.read_buffer(A.BUF_ACTIONS)
    shifted = np.concatenate([mean[1:], mean[-1:]])
    lo, hi = np.asarray(cfg.mppi.u_min, np.float32), np.asarray(cfg.mppi.u_max, np.float32)
    expect = np.clip(shifted[None] + delta * planner.scale_tril.numpy()[None, None], lo, hi)
    assert_close(act[: K - 1], expect[: K - 1], 2e-4, 2e-5, f"{env} perturbation with the adapted scale")   # row K-1: null action
    o.close()


def _linear_world(K, nx_used=4):
    """A tiny differentiable world for the callback paths: state [K,4] = (x, vx, y, vy), u = (ax, ay) (only the first
    two action dimensions act), cost = distance to (1, -1) + 0.1 |u|^2."""
    def dynamics(state, u, t=None):
        s = state.clone()
        s[:, 1] = 0.9 * s[:, 1] + 0.05 * u[:, 0]
        s[:, 3] = 0.9 * s[:, 3] + 0.05 * u[:, 1]
        s[:, 0] = s[:, 0] + 0.05 * s[:, 1]
        s[:, 2] = s[:, 2] + 0.05 * s[:, 3]
        dynamics.last_u = u
        return s, u

    def running_cost(state):
        return torch.sqrt((state[:, 0] - 1.0) ** 2 + (state[:, 2] + 1.0) ** 2) + 0.1 * (dynamics.last_u[:, :2] ** 2).sum(1)
    return dynamics, running_cost


@pytest.mark.parametrize("abs_cost", [False, True])
def test_simple_mode_matches_reference(ref, monkeypatch, abs_cost):
    """mppi_mode='simple' (mppi.py:220-233,335-363): classic MPPI with fresh Gaussian noise, lambda * U Sigma^-1 eps
    control cost, softmin of the total cost at temperature lambda, U += sum_k w_k eps_k. The host mirror (rollouts
    through the callbacks, softmin + weighted sum through m3p2i_update_only -- here on the oracle backend) against the
    imported reference planner fed the same noise, five successive commands."""
    G, ref_m3p2i, ref_cost = ref
    from m3p2i_aip.planners.motion_planner import mppi as our_mppi
    from m3p2i_b200 import native
    monkeypatch.setattr(native, "NativePlanner", O.Oracle)   # the CPU tier has no GPU: the update runs on the oracle
    K, T = 96, 12
    cfg = S.make_cfg("point_env", "navigation", [1.0, -1.0], K, T, device="cpu")
    cfg.mppi.mppi_mode = "simple"
    cfg.mppi.filter_u = False
    cfg.mppi.noise_abs_cost = abs_cost
    cfg.mppi.lambda_ = 0.7
    cfg.mppi.u_per_command = T
    cfg.mppi.noise_sigma = [[2.0, 0.3], [0.3, 1.0]]
    dyn_r, cost_r = _linear_world(K)
    dyn_o, cost_o = _linear_world(K)
    torch.manual_seed(1)
    theirs = ref_m3p2i.M3P2I(cfg, dynamics=dyn_r, running_cost=cost_r)
    ours = our_mppi.MPPI(cfg, dynamics=dyn_o, running_cost=cost_o)
    assert not ours.fused
    U0 = torch.randn(T, 2) * 0.3
    theirs.U, ours.U = U0.clone(), U0.clone()
    rng = np.random.default_rng(5)
    state = torch.tensor([0.0, 0.0, 0.0, 0.0])
    for tick in range(5):
        noise = torch.from_numpy((rng.standard_normal((K, T, 2)) @ np.linalg.cholesky(np.array(cfg.mppi.noise_sigma)).T).astype(np.float32))
        theirs.noise_dist = type("Fixed", (), {"sample": staticmethod(lambda shape, n=noise: n.clone())})()
        ours._sample_noise = lambda *shape, n=noise: n.clone()
        a_t = theirs.command(state)
        a_o = ours.command(state)
        assert_close(a_o.numpy(), a_t.numpy(), 3e-4, 3e-5, f"simple mode action [{tick}]")
        assert_close(ours.U.numpy(), theirs.U.numpy(), 3e-4, 3e-5, f"simple mode U [{tick}]")
        assert_close(ours.weights.numpy(), theirs.weights.numpy(), 3e-3, 1e-7, f"simple mode weights [{tick}]")
        assert_close(ours.cost_total.numpy(), theirs.cost_total.numpy(), 2e-4, 2e-4, f"simple mode cost_total [{tick}]")
        state = torch.tensor([0.02 * (tick + 1), 0.1, -0.01 * (tick + 1), -0.1])
