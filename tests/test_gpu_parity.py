"""CUDA path (through the C ABI, libm3p2i_b200.so) against the CPU oracle and the reference goldens.

Tolerances (fp32; the GPU fuses multiply-adds and uses its own sinf/cosf/expf): contact-free quantities
rtol/atol 1e-4; per-sample costs of contact-rich rollouts may flip the 1000-cost collision threshold
(cost_functions.py:165-169) for <= 0.5 % of the samples; the optimal action is compared with atol 1e-2.
"""
import numpy as np
import pytest

import oracle_py as O
from helpers import assert_close, case_cfg, golden_cases, load_golden, make_backend, tick
from m3p2i_b200 import _abi as A
from m3p2i_b200 import native
from m3p2i_b200 import scene as S

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-3


def _planner_seq(st, key, T, nu):
    return np.asarray(getattr(st, key)[: T * nu], np.float32).reshape(T, nu)


@pytest.mark.parametrize("name", golden_cases())
def test_native_reproduces_reference_goldens(name):
    g = load_golden(name)
    cfg = case_cfg(g)
    n = make_backend(native.NativePlanner, cfg)
    n.set_noise_table(g["delta"])
    mm = bool(g["multi_modal"])
    K, T, nu = int(g["K"]), int(g["T"]), int(g["nu"])
    contact = name in ("nav_obstacle", "push_k256_t20", "pull_k256_t20", "push_pull_mm", "panda_pick")
    bad = 0.02 if contact else 0.0
    if name == "panda_pick":
        bad = 0.01   # grasp state: accumulated + warm-started impulses keep fp32 reordering from growing (round 1: 5 %)
    for i in range(int(g["calls"])):
        action, cost_total, info = tick(n, g, i)
        st = n.get_planner_state()
        assert_close(n.read_buffer(A.BUF_ACTIONS), g[f"actions_{i}"], RTOL, ATOL, f"{name}[{i}] actions")
        assert_close(n.read_buffer(A.BUF_STATES), g[f"states_{i}"], RTOL, ATOL, f"{name}[{i}] states", bad)
        ct, ct_ref = np.asarray(cost_total, np.float64), np.asarray(g[f"cost_total_{i}"], np.float64)
        if contact:
            # cost_total_k = S_k + mean_k(S) (the aliasing quirk, mppi.py:325): one sample whose contact force sits on
            # the 0.1 N threshold of the 1000-cost (cost_functions.py:165-169) would shift EVERY element through the
            # mean. Compare the per-sample sums S_k = cost_total_k - mean(cost_total) / 2, with the flip budget.
            ct, ct_ref = ct - ct.mean() / 2, ct_ref - ct_ref.mean() / 2
        assert_close(ct, ct_ref, RTOL, 5e-3, f"{name}[{i}] cost_total", bad)
        w = n.read_buffer(A.BUF_WEIGHTS)
        assert_close(w[0], g[f"weights_{i}"], 2e-2, 1e-5, f"{name}[{i}] weights", bad)
        assert_close(_planner_seq(st, "mean_action", T, nu), g[f"mean_action_{i}"], 1e-2, 1e-2, f"{name}[{i}] mean_action")
        assert_close(action, g[f"action_{i}"], 1e-2, 1e-2, f"{name}[{i}] action")
        if mm:
            assert int(info.weight_pull > info.weight_push) == int(g[f"pull_preference_{i}"])
        elif str(g["env"]) == "panda_env":
            assert st.beta == pytest.approx(float(g[f"beta_{i}"]), rel=1e-12)
    n.close()


CASES = [
    # name, env, task, goal, K, T, multi_modal, shelf, robot start
    ("nav_c1", "point_env", "navigation", [-3.0, 3.0], 200, 12, False, False, None),
    ("push_c2", "point_env", "push", [-1.0, -1.0], 1024, 20, False, False, [0.2, 2.45]),
    ("push_pull_c3", "point_env", "push_pull", [-3.75, -3.75], 4096, 20, True, False, [0.3, 2.4]),
    ("pull", "point_env", "pull", [0.0, 0.0], 512, 20, False, False, [0.0, 1.55]),
    ("reach", "panda_env", "reach", None, 512, 16, False, False, None),
    ("reach_mm", "panda_env", "reach", None, 512, 16, True, True, None),
    ("pick", "panda_env", "pick", None, 512, 16, False, False, None),
]


def _setup(case, noise_mode, seed=7):
    name, env, task, goal, K, T, mm, shelf, robot = case
    cfg = S.make_cfg(env, task, goal, K, T, multi_modal=mm, cube_on_shelf=shelf)
    actors = S.default_actors(env)
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, shelf).copy()
    if robot:
        dof[0], dof[2] = robot
    if env == "panda_env":
        cb = root[S.actor_index(actors, "cubeB")]
        goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]) if task == "pick" else np.zeros(7)
        # let the cubes rest on the table / shelf as in a running episode
        root[S.actor_index(actors, "cubeA"), 2] -= 0.0095 if not shelf else 0.0
        root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    grip = {"reach": "open", "place": "open", "pick": "close"}.get(task)
    o = make_backend(O.Oracle, cfg, noise_mode=noise_mode, seed=seed)
    n = make_backend(native.NativePlanner, cfg, noise_mode=noise_mode, seed=seed)
    for b in (o, n):
        b.set_state(dof, root)
        b.set_objective(task, goal, grip)
    return cfg, o, n


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_command_matches_oracle_philox(case):
    """Three successive ticks with in-kernel Philox noise against the oracle evaluating the same counters."""
    O.set_threads(8)
    cfg, o, n = _setup(case, A.NOISE_PHILOX)
    K, T, nu = n.K, n.T, n.nu
    assert_close(n.get_noise(), o.get_noise(), 1e-4, 2e-5, f"{case[0]} philox noise")
    contact = case[2] in ("push", "pull", "push_pull", "pick")
    bad = 0.005 if contact else 0.0
    for i in range(3):
        a_o, c_o, i_o = o.command()
        a_n, c_n, i_n = n.command()
        assert_close(n.read_buffer(A.BUF_ACTIONS), o.read_buffer(A.BUF_ACTIONS), 1e-4, 1e-4, f"{case[0]}[{i}] actions", bad)
        assert_close(n.read_buffer(A.BUF_STATES), o.read_buffer(A.BUF_STATES), RTOL, ATOL, f"{case[0]}[{i}] states", bad)
        assert_close(n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON), RTOL, ATOL,
                     f"{case[0]}[{i}] cost_horizon", bad)
        assert_close(c_n, c_o, RTOL, 5e-3, f"{case[0]}[{i}] cost_total", bad)
        assert_close(a_n, a_o, 1e-2, 1e-2, f"{case[0]}[{i}] action")
        assert i_n.launches >= 3
    o.close()
    n.close()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_open_loop_rollout_matches_oracle(case):
    O.set_threads(8)
    cfg, o, n = _setup(case, A.NOISE_TABLE)
    rng = np.random.default_rng(3)
    lo, hi = np.asarray(cfg.mppi.u_min, np.float32), np.asarray(cfg.mppi.u_max, np.float32)
    acts = rng.uniform(lo, hi, size=(n.K, n.T, n.nu)).astype(np.float32)
    s_o, c_o = o.rollout_actions(acts)
    s_n, c_n = n.rollout_actions(acts)
    contact = case[2] in ("push", "pull", "push_pull", "pick")
    bad = 0.005 if contact else 0.0
    assert_close(s_n, s_o, RTOL, ATOL, f"{case[0]} states", bad)
    assert_close(c_n, c_o, RTOL, ATOL, f"{case[0]} cost_horizon", bad)
    o.close()
    n.close()


@pytest.mark.parametrize("env,K,T,mm", [("point_env", 200, 12, False), ("point_env", 4096, 20, True),
                                        ("panda_env", 4096, 32, False), ("panda_env", 1000, 32, True),
                                        ("point_env", 21, 9, False)])
def test_update_only_matches_oracle(env, K, T, mm):
    """Softmin weights + weighted action update on caller-supplied arrays (the generic callback path)."""
    cfg = S.make_cfg(env, "navigation" if env == "point_env" else "reach", None, K, T, multi_modal=mm)
    o = make_backend(O.Oracle, cfg)
    n = make_backend(native.NativePlanner, cfg)
    rng = np.random.default_rng(K + T)
    nu = n.nu
    for rep in range(3):
        ch = (rng.random((K, T)) * 3.0).astype(np.float32)
        ch[rng.integers(0, K, 5)] += 1000.0
        acts = rng.standard_normal((K, T, nu)).astype(np.float32)
        m_o, i_o = o.update_only(ch, acts)
        m_n, i_n = n.update_only(ch, acts)
        assert_close(m_n, m_o, 1e-4, 1e-5, "mean_action")
        assert_close(n.read_buffer(A.BUF_WEIGHTS), o.read_buffer(A.BUF_WEIGHTS), 1e-3, 1e-7, "weights")
        assert list(i_n.best_idx)[: 3 if mm else 1] == list(i_o.best_idx)[: 3 if mm else 1]
        assert_close(list(i_n.eta), list(i_o.eta), 1e-4, 1e-6, "eta")
        assert_close(list(i_n.beta), list(i_o.beta), 1e-6, 0, "beta")
        assert i_n.beta_iters == i_o.beta_iters
        so, sn = o.get_planner_state(), n.get_planner_state()
        assert sn.beta == pytest.approx(so.beta, rel=1e-12)
        for key in ("mean_action", "mean_action_1", "mean_action_2", "best_traj", "best_traj_1", "best_traj_2"):
            assert_close(_planner_seq(sn, key, T, nu), _planner_seq(so, key, T, nu), 1e-4, 1e-5, key)
        idx_o, w_o, _ = o.top_trajs(20)
        idx_n, w_n, _ = n.top_trajs(20)
        assert_close(w_n, w_o, 1e-3, 1e-7, "top weights")
    o.close()
    n.close()


@pytest.mark.parametrize("env,K,T", [("point_env", 512, 20), ("panda_env", 4096, 32)])
def test_update_cov_matches_oracle(env, K, T):
    """mppi.update_cov (mppi.py:508-516) in the update kernels: the variance adapted from the weighted second moment
    (k_wsum's sum w a^2, finish_body) and the noise scale the next fused command perturbs with, against the oracle
    (which tests/test_reference_direct.py pins to the reference's _update_distribution)."""
    task = "navigation" if env == "point_env" else "reach"
    cfg = S.make_cfg(env, task, None, K, T)
    cfg.mppi.update_cov = True
    o = make_backend(O.Oracle, cfg, noise_mode=A.NOISE_PHILOX, seed=3)
    n = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX, seed=3)
    rng = np.random.default_rng(K)
    nu = n.nu
    for rep in range(3):
        ch = (rng.random((K, T)) * 3.0).astype(np.float32)
        acts = rng.uniform(-2, 2, (K, T, nu)).astype(np.float32)
        m_o, _ = o.update_only(ch, acts)
        m_n, _ = n.update_only(ch, acts)
        assert_close(m_n, m_o, 1e-4, 1e-5, "mean_action")
        so, sn = o.get_planner_state(), n.get_planner_state()
        assert_close(np.asarray(sn.cov_action[:nu]), np.asarray(so.cov_action[:nu]), 1e-4, 1e-6, f"cov_action [{rep}]")
        assert not np.allclose(np.asarray(so.cov_action[:nu]), np.asarray(cfg.mppi.noise_sigma).diagonal())
    # the adapted scale drives the next fused command (identical planner state on both sides)
    actors = S.default_actors(env)
    dof, root = S.initial_dof_state(actors), S.initial_root_state(actors)
    n.set_planner_state(o.get_planner_state())
    for b in (o, n):
        b.set_state(dof, root)
        b.set_objective(task, np.zeros(2 if env == "point_env" else 7, np.float32), None)
    a_o, _, _ = o.command()
    a_n, _, _ = n.command()
    assert_close(n.read_buffer(A.BUF_ACTIONS), o.read_buffer(A.BUF_ACTIONS), 1e-4, 1e-4, "actions with the adapted scale")
    assert_close(a_n, a_o, 1e-3, 1e-3, "action")
    so, sn = o.get_planner_state(), n.get_planner_state()
    assert_close(np.asarray(sn.cov_action[:nu]), np.asarray(so.cov_action[:nu]), 1e-4, 1e-6, "cov_action after a fused command")
    o.close()
    n.close()


def test_update_ties_give_nan_like_reference():
    """More than 10 samples tied at the minimum: the multi-modal beta search drives beta to 0 and the reference
    returns NaN weights (m3p2i.py:30-43); the kernel must terminate and agree."""
    K, T = 64, 12
    cfg = S.make_cfg("point_env", "push_pull", None, K, T, multi_modal=True)
    o = make_backend(O.Oracle, cfg)
    n = make_backend(native.NativePlanner, cfg)
    ch = np.ones((K, T), np.float32)
    acts = np.random.default_rng(0).standard_normal((K, T, 2)).astype(np.float32)
    m_o, _ = o.update_only(ch, acts)
    m_n, _ = n.update_only(ch, acts)
    assert np.isnan(m_o).all() and np.isnan(m_n).all()


@pytest.mark.parametrize("env", ["point_env", "panda_env"])
def test_sim_facade_matches_oracle(env):
    K = 64
    cfg = S.sim_only_cfg(env, K, S.make_cfg(env).isaacgym)
    sc = S.build_point_scene() if env == "point_env" else S.build_panda_scene()
    o = O.Oracle(S.build_config(cfg), sc)
    n = native.NativePlanner(S.build_config(cfg), sc)
    actors = S.default_actors(env)
    dof, root = S.initial_dof_state(actors), S.initial_root_state(actors)
    rng = np.random.default_rng(5)
    nu = n.nu
    for b in (o, n):
        b.set_state(dof, root)
    for step in range(25):
        u = rng.uniform(-1.5, 1.5, size=(K, nu)).astype(np.float32)
        if env == "point_env":
            u[:, 1] = np.abs(u[:, 1]) * 2  # drive towards the block at (0, 2)
        for b in (o, n):
            b.sim_set_velocity_target(u)
            b.sim_step()
    for a, b_, what in zip(n.sim_read(), o.sim_read(), ("dof", "root", "link", "contact")):
        assert_close(a, b_, 2e-3, 2e-3, f"{env} sim {what}", 0.01)
    o.close()
    n.close()


def test_panda_fk_known_answers():
    """finger frames at the initial joint pose, derived from franka_panda.urdf."""
    cfg = S.sim_only_cfg("panda_env", 1, S.make_cfg("panda_env").isaacgym)
    n = native.NativePlanner(S.build_config(cfg), S.build_panda_scene())
    actors = S.default_actors("panda_env")
    n.set_state(S.initial_dof_state(actors), S.initial_root_state(actors))
    _, _, link, _ = n.sim_read()
    assert_close(link[0, 0, :3], [0.09540, -0.01414, 1.51177], 0, 2e-5, "left finger")
    assert_close(link[0, 1, :3], [0.06736, 0.01414, 1.51551], 0, 2e-5, "right finger")
    q = link[0, 0, 3:7]
    ref = np.array([-0.92185, -0.38184, 0.06116, 0.02533])
    assert min(np.abs(q - ref).max(), np.abs(q + ref).max()) < 2e-5
    n.close()


def _sample_mismatch(a, b, rtol, atol):
    """fraction of samples (rows) with any element outside the tolerance"""
    bad = ~np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    return bad.reshape(bad.shape[0], -1).any(axis=1).mean()


@pytest.mark.parametrize("case", [CASES[1], CASES[4], CASES[5]], ids=["push_c2", "reach", "reach_mm"])
def test_command_matches_oracle_philox_spline(case):
    """M3P2I_NOISE_PHILOX_SPLINE: the smooth in-kernel noise (B-spline over Philox control points) against the oracle
    evaluating the same counters, thread-per-sample (point) and team kernels (panda), two ticks."""
    O.set_threads(8)
    cfg, o, n = _setup(case, A.NOISE_PHILOX_SPLINE)
    assert_close(n.get_noise(), o.get_noise(), 1e-4, 2e-5, f"{case[0]} philox-spline noise")
    bad = 0.005 if case[2] == "push" else 0.0
    for i in range(2):
        a_o, c_o, _ = o.command()
        a_n, c_n, _ = n.command()
        assert_close(n.read_buffer(A.BUF_ACTIONS), o.read_buffer(A.BUF_ACTIONS), 1e-4, 1e-4, f"{case[0]}[{i}] actions", bad)
        assert_close(n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON), RTOL, ATOL,
                     f"{case[0]}[{i}] cost_horizon", bad)
        assert_close(a_n, a_o, 1e-3, 1e-3, f"{case[0]}[{i}] action")
    o.close()
    n.close()


GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027]  # make_golden.grasp_pose


@pytest.mark.parametrize("task,mm,shelf", [("pick", False, False), ("reach", True, True), ("reach", False, False)])
def test_team_and_thread_kernels_agree(task, mm, shelf):
    """The lane-cooperative (8 and 16 lanes per sample) and the thread-per-sample rollout kernels apply the same impulses
    in the same order (only the summation order of the reported contact forces differs), and both follow the oracle.
    `pick` starts with the fingers closed around cubeA (finger / cube / table contacts in every rollout). Stick /
    slip contact dynamics amplify fp32 rounding differences (FMA contraction, SFU division); with the accumulated,
    warm-started impulse solver 0.2 - 0.6 % of the samples deviate by more than 1e-3 somewhere in a rollout on B200 and
    none flips the collision cost (tests/experiments/grasp_parity.py; the non-accumulated solver of round 1 needed a
    15 % budget); the test allows 2 % and bounds the median."""
    O.set_threads(8)
    case = ("x", "panda_env", task, None, 512, 16, mm, shelf, None)
    budget = 0.02 if task == "pick" else 0.005
    res = {}
    for lanes in (1, 8, 16):
        cfg, o, n = _setup(case, A.NOISE_PHILOX)
        n.close()
        cfg.mppi.lanes_per_sample = lanes
        n = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX, seed=7)
        actors = S.default_actors("panda_env")
        dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, shelf).copy()
        if task == "pick":
            dof[0::2] = GRASP_Q
            root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
            root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
        cb = root[S.actor_index(actors, "cubeB")]
        goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]) if task == "pick" else np.zeros(7)
        for b in (o, n):
            b.set_state(dof, root)
            b.set_objective(task, goal, {"pick": "close", "reach": "open"}[task])
        outs = []
        for i in range(3):
            # every tick starts from the SAME planner state (the oracle's): what is compared is one command on equal
            # inputs; the drift of two closed loops over several ticks is not a property of a kernel
            n.set_planner_state(o.get_planner_state())
            a_n, _, _ = n.command()
            a_o, _, _ = o.command()
            st_n, ch_n = n.read_buffer(A.BUF_STATES), n.read_buffer(A.BUF_COST_HORIZON)
            outs.append((a_n.copy(), st_n, ch_n))
            ch_o = o.read_buffer(A.BUF_COST_HORIZON)
            frac = _sample_mismatch(ch_n, ch_o, RTOL, ATOL)
            assert frac <= budget, f"lanes={lanes} tick {i}: {frac:.3f} of the samples differ from the oracle"
            assert np.median(np.abs(ch_n - ch_o)) < 5e-4
            assert_close(a_n, a_o, 2e-2, 2e-2, f"lanes={lanes} vs oracle action [{i}]")
        res[lanes] = outs
        o.close()
        n.close()
    for team in (8, 16):
        for i in range(3):
            assert_close(res[team][i][0], res[1][i][0], 2e-2, 2e-2, f"team{team} vs thread action [{i}]")
            for j, what in ((1, "states"), (2, "cost_horizon")):
                frac = _sample_mismatch(res[team][i][j], res[1][i][j], RTOL, ATOL)
                assert frac <= budget, f"team{team} vs thread {what} [{i}]: {frac:.3f} of the samples differ"
                assert np.median(np.abs(res[team][i][j] - res[1][i][j])) < 5e-4


@pytest.mark.parametrize("env,task,K,T,lanes", [
    ("panda_env", "pick", 21, 9, 0), ("panda_env", "reach", 20, 64, 0), ("point_env", "push", 33, 64, 0),
    ("point_env", "navigation", 20, 9, 0), ("panda_env", "place", 4609, 12, 0), ("panda_env", "place", 8289, 9, 0), ("panda_env", "place", 12433, 9, 0),
    ("panda_env", "pick", 21, 9, 8), ("panda_env", "reach", 23, 12, 8), ("panda_env", "pick", 2073, 9, 0),
    ("panda_env", "reach", 4145, 9, 0)])
def test_edge_sizes(env, task, K, T, lanes):
    """Ragged sizes: K below one warp / not a multiple of a team, a warp or a CTA for every rollout kernel shape (16
    lanes up to K = 2072, 8 lanes in one / two / three waves of 4144 samples up to 12432, thread per sample beyond;
    lanes = 8 forced at tiny K), the minimum K (20, top-k) and T (9, Savitzky-Golay window), the maximum horizon (64)."""
    O.set_threads(8)
    case = ("edge", env, task, [-1.0, -1.0] if env == "point_env" else None, K, T, False, False, [0.2, 2.45] if task == "push" else None)
    cfg, o, n = _setup(case, A.NOISE_PHILOX)
    if lanes:
        n.close()
        cfg.mppi.lanes_per_sample = lanes
        n = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX, seed=7)
        actors = S.default_actors(env)
        dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, False).copy()
        for b in (o, n):
            b.set_state(dof, root)
            b.set_objective(task, np.zeros(7, np.float32), {"pick": "close", "reach": "open"}[task])
    for i in range(2):
        a_o, c_o, _ = o.command()
        a_n, c_n, info = n.command()
        assert_close(n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON), RTOL, ATOL, f"cost_horizon [{i}]", 0.01)
        assert_close(a_n, a_o, 1e-2, 1e-2, f"action [{i}]")
        assert np.isfinite(c_n).all()
        if env == "panda_env":
            assert info.rollout_lanes == (lanes or (16 if K <= 2072 else 8 if K <= 12432 else 1))
    o.close()
    n.close()
