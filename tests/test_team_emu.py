"""The rollout kernels' DEVICE CODE against the oracle without a GPU: tests/emu compiles panda_team.cuh (the body of
k_rollout_team) and panda_env.cuh (panda_step / panda_cost, the thread-per-sample path) for the host and runs the
team kernel 32 lanes in lock step (fibers; every warp collective checks that all lanes meet at the same source line,
so a non-uniform branch around a shuffle -- undefined behaviour on the GPU -- fails here). Open-loop rollouts of the
same actions from the same state must reproduce the oracle's costs and state rows: rest state (both cubes asleep),
fingers closing on cubeA and carrying it (accumulated impulses, warm start, cube / table contact), the pre-grasp
straddle, reach with the producer rows (single- and multi-modal, cube on the shelf), place."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

import oracle_py as O  # noqa: E402
import team_emu as E  # noqa: E402
from m3p2i_b200 import _abi as A  # noqa: E402
from m3p2i_b200 import scene as S  # noqa: E402

GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333]


def _case(task, fingers, K, T, shelf=False, mm=False, seed=0, sigma=1.0, lift=0.0):
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm, cube_on_shelf=shelf)
    c = S.build_config(cfg, noise_mode=A.NOISE_TABLE, seed=0)
    scene = S.build_panda_scene()
    actors = S.default_actors("panda_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors, shelf).copy()
    if not shelf:
        root[S.actor_index(actors, "cubeA"), 2] -= 0.0095
    root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
    if fingers is not None:
        dof[0::2] = GRASP_Q + [fingers, fingers]
        dof[2] -= lift   # shoulder back: the gripper starts `lift` rad above the grasp pose

    cb = root[S.actor_index(actors, "cubeB")]
    goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]]).astype(np.float32) if task == "pick" \
        else np.zeros(7, np.float32)
    rng = np.random.default_rng(seed)
    a = np.clip(rng.normal(size=(K, 1, 9)) * sigma + 0.3 * sigma * rng.normal(size=(K, T, 9)), -2, 2).astype(np.float32)
    grip = {"pick": "close", "reach": "open", "place": "open"}[task]
    a[:, :, 7:] = -1.5 if grip == "close" else 1.5
    o = O.Oracle(c, scene)
    o.set_state(dof, root)
    o.set_objective(task, goal, grip)
    st_o, ch_o = o.rollout_actions(a)
    o.close()
    return c, scene, task, goal, grip, dof, root, a, st_o, ch_o


@pytest.mark.parametrize("task,fingers,K,T,lanes,shelf,mm,sigma", [
    ("pick", None, 12, 6, 8, False, False, 1.0),     # arm away from the cubes: both asleep, bit-equal costs
    ("pick", 0.027, 16, 10, 8, False, False, 1.0),   # closed on cubeA: squeeze, lift, table contact
    ("pick", 0.027, 8, 10, 16, False, False, 1.0),
    ("pick", 0.04, 16, 10, 8, False, False, 0.5),    # open fingers astride cubeA closing on it
    ("reach", 0.04, 12, 8, 16, False, False, 0.5),
    ("reach", None, 16, 6, 8, True, True, 1.0),
    ("place", 0.027, 8, 9, 8, False, False, 1.0),
])
def test_team_device_code_matches_oracle(task, fingers, K, T, lanes, shelf, mm, sigma):
    c, scene, task, goal, grip, dof, root, a, st_o, ch_o = _case(task, fingers, K, T, shelf, mm, sigma=sigma)
    st_e, ch_e, env_e, ncoll = E.rollout_actions(c, scene, task, goal, grip, dof, root, a, lanes)
    assert ncoll > 0
    assert np.allclose(st_e, st_o, rtol=1e-5, atol=1e-5)
    bad = ~np.isclose(ch_e, ch_o, rtol=1e-3, atol=5e-3)
    assert not bad.any(), f"{bad.sum()} of {bad.size} step costs differ, max {np.abs(ch_e - ch_o).max()}"
    if fingers is None and task == "pick":
        assert np.array_equal(ch_e, ch_o)   # nothing but the arm moves
    if task != "reach":
        st_t, ch_t, env_t = E.thread_rollout_actions(c, scene, task, goal, grip, dof, root, a)
        assert np.allclose(st_t, st_o, rtol=1e-5, atol=1e-5)
        assert np.isclose(ch_t, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch_t - ch_o).max()
        assert np.abs(env_t[:, :44] - env_e[:, :44]).max() < 5e-3   # end states of the two kernel shapes


@pytest.mark.parametrize("lift,lanes", [(0.15, 8), (0.15, 16), (0.3, 8)])
def test_dormant_cubes_wake_up_like_the_oracle(lift, lanes):
    """The gripper starts 0.2 - 0.3 m above cubeA -- inside / just outside the bounding sphere of hand + fingers the dormant
    shortcut tests first -- and random actions take some rollouts down onto the cube and others away from it: the
    sub-steps the kernels skip (both cubes asleep, no link box in reach) and the ones they do not must add up to the
    oracle's trajectory, which has no such shortcut."""
    c, scene, task, goal, grip, dof, root, a, st_o, ch_o = _case("pick", 0.04, 16, 12, sigma=1.5, lift=lift, seed=3)
    st_e, ch_e, env_e, _ = E.rollout_actions(c, scene, task, goal, grip, dof, root, a, lanes)
    assert np.allclose(st_e, st_o, rtol=1e-5, atol=1e-5)
    assert np.isclose(ch_e, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch_e - ch_o).max()
    st_t, ch_t, env_t = E.thread_rollout_actions(c, scene, task, goal, grip, dof, root, a)
    assert np.isclose(ch_t, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch_t - ch_o).max()
    moved = np.abs(env_e[:, 18:21] - env_e[0:1, 18:21]).max(axis=1) > 1e-4   # cubeA ended somewhere else than in rollout 0
    if lift < 0.2:
        assert moved.any() and not moved.all()   # some rollouts reached the cube, some never woke it


def test_contact_rich_case_has_contacts():
    """The grasp case above is not vacuous: the oracle's rollouts contain collision-cost steps and the cube moves."""
    *_, st_o, ch_o = _case("pick", 0.027, 16, 10)
    assert (ch_o > 900).any() and (ch_o < 900).any()


@pytest.mark.parametrize("task,fingers,K,T,shelf,mm,lift,expect", [
    ("pick", None, 12, 32, False, False, 0.0, "most"),     # the bench state: arm at its initial pose, cubes asleep
    ("reach", None, 16, 32, True, True, 0.0, "most"),      # C5's state: multi-modal reach, cube on the shelf
    ("place", None, 8, 12, False, False, 0.0, "most"),
    ("pick", 0.04, 32, 24, False, False, 0.5, "some"),     # gripper 0.3 m above cubeA: some rollouts come down to it
    ("pick", 0.027, 8, 10, False, False, 0.0, "none"),     # fingers closed on cubeA: nothing is far
])
def test_far_field_code_matches_team_and_oracle(task, fingers, K, T, shelf, mm, lift, expect):
    """panda_far.cuh (the body of k_rollout_far): the samples it declares far carry exactly the costs, sums and state rows
    the team kernel's device code and the oracle produce for them, and their cubes indeed never move; the others are left
    to the full rollout."""
    c, scene, task, goal, grip, dof, root, a, st_o, ch_o = _case(task, fingers, K, T, shelf, mm, lift=lift)
    ok, pok, st, ch, cs, J = E.far_rollout_actions(c, scene, task, goal, grip, dof, root, a)
    st_e, ch_e, env_e, _ = E.rollout_actions(c, scene, task, goal, grip, dof, root, a, 8)
    assert {"most": ok.sum() >= K - 2, "some": 0 < ok.sum() < K, "none": ok.sum() == 0}[expect], ok.sum()
    if expect != "none":                                  # (gripper next to a cube at the start: early-out, nothing stored)
        assert np.array_equal(st, st_e)                  # state rows: the same joint recurrences
    assert np.array_equal(ch[ok], ch_e[ok])              # costs: the same functions on the same poses
    assert np.isclose(ch[ok], ch_o[ok], rtol=1e-6, atol=2e-6).all()
    if ok.any():
        actors = S.default_actors("panda_env")
        ra, rb = (np.asarray(root).reshape(-1, 13)[S.actor_index(actors, n)] for n in ("cubeA", "cubeB"))
        start = np.concatenate([ra[:7], np.zeros(6, np.float32), rb[:7], np.zeros(6, np.float32)])
        assert np.abs(env_e[ok][:, 18:44] - start).max() < 1e-6   # far samples: both cubes asleep where they started
        g = np.float32(c.gamma) ** np.arange(T, dtype=np.float32)
        assert np.allclose(cs[ok], ch[ok].sum(1), rtol=1e-5) and np.allclose(J[ok], (ch[ok] * g).sum(1), rtol=1e-5)
    if task == "reach":
        assert pok.all()   # rows 0 and K/2 stay far: the batch rows are the start pose


@pytest.mark.parametrize("task,fingers,K,T,lift,sigma,seed", [
    ("pick", None, 12, 32, 0.0, 1.0, 0),      # bench state: one rollout wanders off towards the table late in the horizon
    ("pick", 0.04, 32, 24, 0.5, 1.0, 0),      # hand-overs at boundaries 0, 1, 2 and 4
    ("pick", 0.04, 32, 32, 0.6, 1.0, 1),
    ("place", 0.04, 16, 24, 0.5, 1.0, 0),
    ("pick", 0.04, 16, 12, 0.15, 1.5, 3),     # the gripper starts inside its bounding sphere (exact tests decide), early hand-overs
    ("pick", 0.04, 24, 32, 0.45, 1.5, 2),
])
@pytest.mark.parametrize("lanes", [8, 16])
def test_far_split_with_hand_over_matches_oracle(task, fingers, K, T, lift, sigma, seed, lanes):
    _split_case(task, fingers, K, T, False, False, lift, sigma, seed, lanes)


@pytest.mark.parametrize("fingers,K,T,shelf,mm,lift,sigma,seed", [
    (None, 16, 32, True, True, 0.0, 1.0, 0),       # C5's state: rows 0 and K/2 stay far, published by the far-field code
    (0.04, 32, 24, False, False, 0.5, 1.0, 0),     # near samples wait for nothing: the rows are already there
    (0.04, 32, 24, False, True, 0.5, 1.0, 1),
    (0.04, 32, 32, False, False, 0.45, 1.5, 2),    # hand-overs at 3, 5, 6, 7 with deferred reach costs
    (0.04, 12, 8, False, False, 0.0, 0.5, 0),      # gripper astride cubeA: early-out, the producer CTA replays the rows
])
@pytest.mark.parametrize("lanes", [8, 16])
def test_far_split_reach_matches_oracle(fingers, K, T, shelf, mm, lift, sigma, seed, lanes):
    _split_case("reach", fingers, K, T, shelf, mm, lift, sigma, seed, lanes)


def _split_case(task, fingers, K, T, shelf, mm, lift, sigma, seed, lanes):
    """What a pick / place command launches: the far-field code over all samples, then the team kernel over the near list,
    every listed sample starting at its warp's hand-over boundary (joints from the dump, finished costs from cost_h).
    All K samples must carry the oracle's costs, sums and state rows."""
    c, scene, task, goal, grip, dof, root, a, st_o, ch_o = _case(task, fingers, K, T, shelf, mm, lift=lift, sigma=sigma, seed=seed)
    st, ch, cs, J, far, bd = E.split_rollout_actions(c, scene, task, goal, grip, dof, root, a, lanes)
    assert np.allclose(st, st_o, rtol=1e-5, atol=1e-5)
    assert np.isclose(ch, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch - ch_o).max()
    g = np.float32(c.gamma) ** np.arange(T, dtype=np.float32)
    assert np.allclose(cs, ch.sum(1), rtol=1e-5, atol=1e-4) and np.allclose(J, (ch * g).sum(1), rtol=1e-5, atol=1e-4)
    if task == "pick" and lift in (0.5, 0.6):
        assert far.any() and (~far).any() and (bd[~far] > 0).any()   # far samples, near samples and real hand-overs


@pytest.mark.parametrize("task,T,ns", [("pick", 32, 1), ("pick", 32, 4), ("reach", 32, 4), ("pick", 36, 1), ("pick", 30, 2)])
@pytest.mark.parametrize("lanes", [8, 16])
def test_far_split_other_substeps_and_horizons(task, T, ns, lanes):
    """Hand-over boundaries are multiples of 8 ITERATIONS: 1 and 4 sub-steps per step and horizons whose iteration count
    is not a multiple of the run-ahead block (16) must hand over at step boundaries all the same."""
    c, scene, task, goal, grip, dof, root, a, _, _ = _case(task, 0.04, 24, T, lift=0.45, sigma=1.5, seed=2)
    c.substeps = ns
    o = O.Oracle(c, scene)
    o.set_state(dof, root)
    o.set_objective(task, goal, grip)
    st_o, ch_o = o.rollout_actions(a)
    o.close()
    st, ch, cs, J, far, bd = E.split_rollout_actions(c, scene, task, goal, grip, dof, root, a, lanes)
    assert far.any() and (~far).any() and (bd[~far] > 0).all()
    assert np.allclose(st, st_o, rtol=1e-5, atol=1e-5)
    assert np.isclose(ch, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch - ch_o).max()
    assert np.allclose(cs, ch.sum(1), rtol=1e-5, atol=1e-4)


def _base_mods():
    actors = S.default_actors("panda_env")
    ia, ib = S.actor_index(actors, "cubeA"), S.actor_index(actors, "cubeB")

    def lift(r): r[ia, 2] += 0.02
    def moving(r): r[ia, 7] = 0.02
    def creeping(r): r[ia, 7] = 0.003
    def spinning(r): r[ia, 12] = 0.2
    def tilted(r): r[ia, 3:7] = [np.sin(0.15), 0, 0, np.cos(0.15)]; r[ia, 2] += 0.01
    def stacked(r): r[ia, :3] = r[ib, :3] + np.array([0, 0, 0.05], np.float32)
    def touching(r): r[ia, :3] = r[ib, :3] + np.array([0.0505, 0, 0], np.float32)
    def edge(r): r[ia, 0] = 0.58
    def both(r): r[ia, 2] += 0.02; r[ib, 2] += 0.03
    return {"rest": (lambda r: None, True), "lift": (lift, False), "moving": (moving, False), "creeping": (creeping, True),
            "spinning": (spinning, False), "tilted": (tilted, False), "stacked": (stacked, False), "touching": (touching, False),
            "edge": (edge, False), "both": (both, False)}


@pytest.mark.parametrize("name", ["rest", "lift", "moving", "creeping", "spinning", "tilted", "stacked", "touching", "edge", "both"])
def test_far_field_start_state_rule(name):
    """far_base_asleep: the far-field code may only take a command whose cubes the full path would put to sleep in the very
    first sub-step (at rest on their support, below the speed thresholds, not touching each other). Start states with a cube
    in the air, moving, tilted, stacked, touching the other cube or hanging over the table edge must all go to the rollout
    kernel -- and whatever is decided, every sample carries the oracle's costs."""
    mod, expect_far = _base_mods()[name]
    c, scene, task, goal, grip, dof, root, a, _, _ = _case("pick", None, 8, 16)
    root = np.asarray(root, np.float32).reshape(-1, 13).copy()
    mod(root)
    o = O.Oracle(c, scene)
    o.set_state(dof, root)
    o.set_objective(task, goal, grip)
    st_o, ch_o = o.rollout_actions(a)
    o.close()
    st, ch, cs, J, far, bd = E.split_rollout_actions(c, scene, task, goal, grip, dof, root, a, 8)
    assert far.all() if expect_far else not far.any(), far
    assert np.allclose(st, st_o, rtol=1e-5, atol=1e-5)
    assert np.isclose(ch, ch_o, rtol=1e-3, atol=5e-3).all(), np.abs(ch - ch_o).max()
