"""Size-independent properties of the CUDA path at the full BASELINE.json sizes (the oracle is too slow there to
be the checker for every element; it still checks a strided subset of samples)."""
import numpy as np
import pytest

import oracle_py as O
from helpers import assert_close, make_backend
from m3p2i_b200 import _abi as A
from m3p2i_b200 import native
from m3p2i_b200 import scene as S

pytestmark = pytest.mark.gpu


def _c4(K=4096, T=32, task="pick", mm=False, noise=A.NOISE_PHILOX, K_local=None, offset=0):
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm)
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors).copy()
    root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    cb = root[S.actor_index(actors, "cubeB")]
    goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]])
    n = make_backend(native.NativePlanner, cfg, noise_mode=noise, seed=11, K_local=K_local, offset=offset)
    n.set_state(dof, root)
    n.set_objective(task, goal, {"pick": "close", "reach": "open"}[task])
    return cfg, n, (dof, root, goal)


def test_c4_invariants():
    cfg, n, _ = _c4()
    K, T, nu = n.K, n.T, n.nu
    action, cost_total, info = n.command()
    acts, states = n.read_buffer(A.BUF_ACTIONS), n.read_buffer(A.BUF_STATES)
    ch, w = n.read_buffer(A.BUF_COST_HORIZON), n.read_buffer(A.BUF_WEIGHTS)
    # bounds, gripper override, null action (mppi.py:300-302,405,412-416)
    lo, hi = np.asarray(cfg.mppi.u_min, np.float32), np.asarray(cfg.mppi.u_max, np.float32)
    assert (acts >= lo - 1e-6).all() and (acts <= hi + 1e-6).all()
    assert (acts[:-1, :, 7:] == -1.5).all()
    assert (acts[-1] == 0).all()
    # cost_total = sum_t c + mean_k sum_t c (mppi.py:282-284,325)
    cs = ch.sum(1, dtype=np.float64)
    assert_close(cost_total, cs + cs.mean(), 1e-5, 1e-3, "cost_total quirk")
    # weights: softmin of the discounted cost (mppi.py:435-442)
    J = n.read_buffer(A.BUF_COST_DISC)
    g = 0.95 ** np.arange(T)
    assert_close(J, (ch.astype(np.float64) * g).sum(1), 1e-5, 1e-4, "discounted cost")
    assert w[0].sum() == pytest.approx(1.0, abs=1e-4)
    assert int(np.argmax(w[0])) == int(np.argmin(J)) == info.best_idx[0]
    e = np.exp(-(J.astype(np.float64) - J.min()) / info.beta[0])
    assert_close(w[0], e / e.sum(), 1e-3, 1e-7, "weights")
    # mean update (mppi.py:498-503), first call: old mean is zero
    new_mean = (w[0][:, None, None].astype(np.float64) * acts).sum(0)
    st = n.get_planner_state()
    mean = np.asarray(st.mean_action[: T * nu], np.float32).reshape(T, nu)
    assert_close(mean, 0.98 * new_mean, 1e-4, 1e-5, "mean_action")
    best = np.asarray(st.best_traj[: T * nu], np.float32).reshape(T, nu)
    assert np.array_equal(best, acts[info.best_idx[0]])
    # filtered action = S @ mean (mppi.py:257-263)
    assert_close(action, S.savgol_matrix(T) @ mean, 1e-4, 1e-5, "savgol")
    # joints stay inside the URDF limits, states are joint 1/2 rows
    assert (np.abs(states[:, :, 1]) <= 2.175 + 1e-5).all() and (np.abs(states[:, :, 3]) <= 2.175 + 1e-5).all()
    assert np.isfinite(cost_total).all()
    n.close()


def test_c4_deterministic_and_shift():
    _, a, _ = _c4()
    _, b, _ = _c4()
    for _ in range(3):
        ra, rb = a.command(), b.command()
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    a.close()
    b.close()


def test_c4_strided_samples_match_oracle():
    """Every 64th sample of the K=4096 x H=32 pick rollout against the oracle run on just those samples."""
    O.set_threads(8)
    cfg, n, (dof, root, goal) = _c4()
    n.command()
    acts = n.read_buffer(A.BUF_ACTIONS)
    ch = n.read_buffer(A.BUF_COST_HORIZON)
    states = n.read_buffer(A.BUF_STATES)
    sel = np.arange(0, 4096, 64)
    ocfg = S.make_cfg("panda_env", "pick", None, len(sel), 32)
    o = make_backend(O.Oracle, ocfg)
    o.set_state(dof, root)
    o.set_objective("pick", goal, "close")
    # open-loop replay of the exact actions the kernel drew (sample K-1 is not in the subset: no null action here)
    ocfg.mppi.sample_null_action = False
    o2 = make_backend(O.Oracle, ocfg)
    o2.set_state(dof, root)
    o2.set_objective("pick", goal, "close")
    s_o, c_o = o2.rollout_actions(acts[sel])
    assert_close(states[sel], s_o, 1e-3, 1e-3, "states", 0.005)
    assert_close(ch[sel], c_o, 1e-3, 1e-3, "cost_horizon", 0.005)
    o.close(); o2.close(); n.close()


@pytest.mark.parametrize("task,mm", [("pick", False), ("reach", True)])
def test_shard_invariance(task, mm):
    """K split over 2 and 4 shards (host-staged phases on one GPU): per-sample costs are bit-identical to the
    unsharded run (Philox counters use the global sample id), the action agrees to reduction order."""
    K, T = 2048, 16
    _, full, _ = _c4(K, T, task, mm)
    a_full, c_full, _ = full.command()
    a_full, c_full = a_full.copy(), c_full.copy()
    J_full = full.read_buffer(A.BUF_COST_DISC)
    for nshard in (2, 4):
        Kl = K // nshard
        shards = [_c4(K, T, task, mm, K_local=Kl, offset=r * Kl)[1] for r in range(nshard)]
        J = np.concatenate([s.phase_rollout() for s in shards])
        assert np.array_equal(J, J_full)
        parts = sum(s.phase_partials(J) for s in shards)
        outs = [s.phase_finish(parts) for s in shards]
        for r, (a, c, info) in enumerate(outs):
            assert_close(a, a_full, 1e-5, 1e-6, f"action shard {r}/{nshard}")
            assert_close(c, c_full[r * Kl:(r + 1) * Kl], 1e-6, 1e-4, f"cost_total shard {r}/{nshard}")
        for s in shards:
            s.close()
    full.close()


@pytest.mark.parametrize("task,mm,nshard", [("pick", False, 2), ("reach", True, 2), ("pick", False, 4)])
def test_peer_memory_exchange_matches_unsharded(task, mm, nshard):
    """The exchange fused into the kernels (m3p2i_peer_export / m3p2i_peer_attach: the rollout kernel stores J into
    every rank's mailbox, the weighted-sum kernel's last CTA pushes and reduces the partial sums) with the ranks as
    handles of ONE process on ONE GPU -- the same kernels and flag protocol as one process per GPU over NVLink, where
    only the mapping of the mailboxes (cudaIpc) differs. Three ticks; every rank must report the unsharded action
    and its slice of the unsharded per-sample costs, and all ranks must agree bit for bit."""
    K, T = 2048, 16
    _, full, _ = _c4(K, T, task, mm)
    Kl = K // nshard
    shards = [_c4(K, T, task, mm, K_local=Kl, offset=r * Kl)[1] for r in range(nshard)]
    desc = [s.peer_export() for s in shards]
    for r, s in enumerate(shards):
        s.peer_attach(r, nshard, desc)
    for tick in range(3):
        a_full, c_full, info_full = full.command()
        a_full, c_full = a_full.copy(), c_full.copy()
        J_full = full.read_buffer(A.BUF_COST_DISC)
        for s in shards:
            s.command_resident()          # asynchronous: every rank's kernels must be in flight before any fetch
        outs = [(a.copy(), c.copy()) for a, c in (s.fetch_result() for s in shards)]
        for r, (a, c) in enumerate(outs):
            assert np.array_equal(shards[r].read_buffer(A.BUF_COST_DISC), J_full), f"tick {tick} rank {r}: gathered J"
            assert_close(a, a_full, 1e-5, 1e-6, f"tick {tick} action rank {r}/{nshard}")
            assert_close(c, c_full[r * Kl:(r + 1) * Kl], 1e-6, 1e-4, f"tick {tick} cost_total rank {r}/{nshard}")
            assert np.array_equal(a, outs[0][0]), f"tick {tick}: rank {r} and rank 0 disagree"
    for s in shards:
        s.close()
    full.close()


def test_peer_exchange_reports_a_missing_rank(monkeypatch):
    """A rank that never delivers must not hang the GPU: the waits are bounded and the fetch fails loudly."""
    monkeypatch.setenv("M3P2I_PEER_TIMEOUT_MS", "300")
    K, T = 256, 12
    shards = [_c4(K, T, "pick", False, K_local=K // 2, offset=r * (K // 2))[1] for r in range(2)]
    desc = [s.peer_export() for s in shards]
    for r, s in enumerate(shards):
        s.peer_attach(r, 2, desc)
    shards[0].command_resident()          # rank 1 never runs
    with pytest.raises(native.NativeError, match="peer exchange timed out"):
        shards[0].fetch_result()
    for s in shards:
        s.close()


def test_table_and_philox_agree():
    """Feeding the dumped Philox table back in table mode gives the same rollout."""
    _, p, _ = _c4(1024, 16)
    _, t, _ = _c4(1024, 16, noise=A.NOISE_TABLE)
    t.set_noise_table(p.get_noise())
    rp, rt = p.command(), t.command()
    assert_close(rt[1], rp[1], 1e-6, 1e-5, "cost_total")
    assert_close(rt[0], rp[0], 1e-6, 1e-6, "action")
    p.close(); t.close()


def test_errors_are_loud():
    cfg = S.make_cfg("point_env", "navigation", [1.0, 1.0], 64, 12)
    c = S.build_config(cfg)
    n = native.NativePlanner(c, S.build_point_scene())
    with pytest.raises(native.NativeError, match="set_state"):
        n.command()
    with pytest.raises(native.NativeError, match="task"):
        n.set_objective("pick", [0.0] * 7)
    bad = S.build_config(cfg)
    bad.nu = 3
    with pytest.raises(native.NativeError, match="nu"):
        native.NativePlanner(bad, S.build_point_scene())
    n.close()


@pytest.mark.parametrize("offset", [0, 16384])
def test_c5_shard_matches_oracle(offset):
    """BASELINE configs[4] as stated: config_panda multi_modal=True cube_on_shelf=True, K_global = 32768, H = 32, sharded
    over 8 GPUs. One GPU's shard (K_local = 4096) at the two interesting places -- the shard that owns global row 0 and
    the one that starts at the mode boundary K/2 = 16384 (it owns the row whose cube axis every second-mode cost reads,
    skill_utils.py:275-279, and must replay row 0 of another shard, cost_functions.py:98) -- against the oracle
    evaluating the same shard with the same Philox counters: actions, per-step costs, discounted costs. A second tick
    follows from planner sequences that are no longer zero (both shards must rebuild the foreign rows from them)."""
    O.set_threads(8)
    Kg, Kl, T = 32768, 4096, 32
    cfg = S.make_cfg("panda_env", "reach", None, Kg, T, multi_modal=True, cube_on_shelf=True)
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, True).copy()
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    pair = []
    for cls in (O.Oracle, native.NativePlanner):
        b = make_backend(cls, cfg, noise_mode=A.NOISE_PHILOX, seed=5, K_local=Kl, offset=offset)
        b.set_state(dof, root)
        b.set_objective("reach", np.zeros(7, np.float32), "open")
        pair.append(b)
    o, n = pair
    rng = np.random.default_rng(3)
    for tick in range(2):
        if tick == 1:
            st = o.get_planner_state()   # non-trivial means / best rows for both modes, identical on both backends
            for key in ("mean_action", "mean_action_1", "mean_action_2", "best_traj_1", "best_traj_2"):
                arr = getattr(st, key)
                for i, v in enumerate(rng.uniform(-0.5, 0.5, T * 9).astype(np.float32)):
                    arr[i] = float(v)
            o.set_planner_state(st)
            n.set_planner_state(st)
        J_o, J_n = o.phase_rollout(), n.phase_rollout()
        assert_close(n.read_buffer(A.BUF_ACTIONS), o.read_buffer(A.BUF_ACTIONS), 1e-4, 1e-4, f"c5 shard@{offset}[{tick}] actions")
        assert_close(n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON), 1e-3, 1e-3,
                     f"c5 shard@{offset}[{tick}] cost_horizon", 0.005)
        assert_close(J_n, J_o, 1e-3, 1e-3, f"c5 shard@{offset}[{tick}] discounted costs", 0.005)
    o.close()
    n.close()


GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333]


@pytest.mark.parametrize("task,mm,shelf,lift,K", [
    ("pick", False, False, None, 4096),    # C4 bench state: every rollout stays in the far field
    ("reach", True, True, None, 4096),     # C5 shard: reach costs read the batch rows
    ("pick", False, False, 0.5, 2048),     # gripper 0.3 m above cubeA: far and near samples mixed (16-lane teams)
    ("pick", False, False, 0.5, 16384),    # the same through the thread-per-sample kernel
    ("reach", False, False, 0.45, 4096),   # reach with near samples: the producer CTA publishes the rows
    ("pick", False, False, 0.0, 1024),     # fingers around cubeA: nothing is far
])
def test_far_field_split_matches_full_rollout(monkeypatch, task, mm, shelf, lift, K):
    """k_rollout_far + the rollout kernel over the near list (panda_far.cuh) against the rollout kernel over all K
    samples (M3P2I_FAR=0): same per-sample costs, same state rows, same command, over several ticks."""
    T = 32
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm, cube_on_shelf=shelf)
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, shelf).copy()
    if not shelf:
        root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    if lift is not None:
        dof[0::2] = GRASP_Q + [0.04, 0.04]
        dof[2] -= lift
    cb = root[S.actor_index(actors, "cubeB")]
    goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]) if task == "pick" else np.zeros(7, np.float32)
    res = {}
    for far in ("1", "0"):
        monkeypatch.setenv("M3P2I_FAR", far)
        n = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX, seed=5)
        n.set_state(dof, root)
        n.set_objective(task, goal, {"pick": "close", "reach": "open"}[task])
        out = []
        for _ in range(3):
            action, cost_total, info = n.command()
            out.append((action, cost_total, n.read_buffer(A.BUF_COST_HORIZON), n.read_buffer(A.BUF_STATES),
                        n.read_buffer(A.BUF_ACTIONS), info.launches))
        res[far] = out
        n.close()
    for i, ((a1, c1, ch1, st1, ac1, l1), (a0, c0, ch0, st0, ac0, l0)) in enumerate(zip(res["1"], res["0"])):
        # one more launch, the far-field kernel; beyond the team kernel's range (K > 84 SMs) two: team AND thread-per-sample
        # kernel are launched over the near list and its length decides on the device which one works
        assert l1 == l0 + (2 if K > 12432 else 1)
        if i == 0:
            assert np.array_equal(ac1, ac0) and np.array_equal(st1, st0)   # same planner state: same actions, same joints
        else:
            # later ticks start from planner states that may differ in the last bits (sums in a different kernel)
            assert np.allclose(ac1, ac0, atol=1e-4) and np.allclose(st1, st0, atol=1e-4)
        bad = ~np.isclose(ch1, ch0, rtol=1e-4, atol=1e-4)
        assert bad.sum() <= 0.002 * bad.size, f"{bad.sum()} of {bad.size} step costs differ, max {np.abs(ch1 - ch0).max()}"
        assert_close(a1, a0, 1e-4, 1e-4, "command")
