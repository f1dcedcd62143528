"""Tiny commands through the far-field split (k_rollout_far -> near list with hand-over boundaries -> rollout kernel),
meant to be run under compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_far.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), ROOT]
from m3p2i_b200 import _abi as A, native, scene as S  # noqa: E402
import bench  # noqa: E402

GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.04, 0.04]


def panda(K, T, task, lift=None, mm=False, shelf=False, lanes=0, K_local=None, offset=0):
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm, cube_on_shelf=shelf)
    cfg.mppi.lanes_per_sample = lanes
    c = S.build_config(cfg, num_samples_local=K_local or K, sample_offset=offset, noise_mode=A.NOISE_PHILOX, seed=1)
    p = native.NativePlanner(c, S.build_panda_scene())
    p.set_filter_matrix(S.savgol_matrix(T))
    dof, root, goal = bench.scene_inputs(dict(env="panda_env", task=task, shelf=shelf))
    if lift is not None:   # gripper `lift` rad of shoulder above the grasp pose: some rollouts come down onto cubeA
        dof = dof.copy()
        dof[0::2] = GRASP_Q
        dof[2] -= lift
    p.set_state(dof, root)
    p.set_objective(task, goal if task == "pick" else np.zeros(7, np.float32), "close" if task == "pick" else "open")
    return p


for name, p in (("pick, all far", panda(64, 32, "pick")),
                ("pick, hand-overs, 16 lanes", panda(96, 32, "pick", lift=0.5)),
                ("pick, hand-overs, 8 lanes", panda(96, 30, "pick", lift=0.45, lanes=8)),
                ("pick, hand-overs, thread per sample", panda(64, 32, "pick", lift=0.45, lanes=1)),
                ("reach mm shelf, rows published by the far kernel", panda(64, 16, "reach", mm=True, shelf=True)),
                ("reach, hand-overs + deferred costs", panda(96, 32, "reach", lift=0.45, lanes=8)),
                ("pick, gripper astride cubeA: early-out", panda(48, 12, "pick", lift=0.0))):
    for _ in range(2):
        a, c, info = p.command()
    assert np.isfinite(a).all() and np.isfinite(c).all(), name
    print("ok:", name, "near", info.near_samples, "launches", info.launches, flush=True)
    p.close()
shards = [panda(64, 16, "pick", lift=0.5, K_local=32, offset=32 * r) for r in range(2)]
desc = [s.peer_export() for s in shards]
for r, s in enumerate(shards):
    s.peer_attach(r, 2, desc)
for _ in range(2):
    for s in shards:
        s.command_resident()
    res = [s.fetch_result() for s in shards]
assert np.array_equal(res[0][0], res[1][0])
print("ok: two shards over peer memory with the far-field split", flush=True)
for s in shards:
    s.close()
