"""A handful of tiny commands covering every rollout kernel shape and the peer-memory exchange, meant to be run under
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), ROOT]
from m3p2i_b200 import _abi as A, native, scene as S  # noqa: E402
import bench  # noqa: E402


GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027]   # fingers on cubeA


def panda(K, T, task, lanes=0, mm=False, K_local=None, offset=0, noise=A.NOISE_PHILOX, grasp=False):
    cfg = S.make_cfg("panda_env", task, None, K, T, multi_modal=mm)
    cfg.mppi.lanes_per_sample = lanes
    c = S.build_config(cfg, num_samples_local=K_local or K, sample_offset=offset, noise_mode=noise, seed=1)
    p = native.NativePlanner(c, S.build_panda_scene())
    p.set_filter_matrix(S.savgol_matrix(T))
    dof, root, goal = bench.scene_inputs()
    if grasp:   # contact-rich: link / cube contact records, accumulators, warm start
        dof = dof.copy()
        dof[0::2] = GRASP_Q
    if noise == A.NOISE_TABLE:
        p.set_noise_halton_spline()   # the table is built on the device (k_halton_spline)
    p.set_state(dof, root)
    p.set_objective(task, goal if task == "pick" else np.zeros(7, np.float32), "close" if task == "pick" else "open")
    return p


def point(K, T, task, goal, mm=False):
    cfg = S.make_cfg("point_env", task, goal, K, T, multi_modal=mm)
    p = native.NativePlanner(S.build_config(cfg, noise_mode=A.NOISE_PHILOX_SPLINE, seed=1), S.build_point_scene())
    p.set_filter_matrix(S.savgol_matrix(T))
    actors = S.default_actors("point_env")
    dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors)
    dof[0], dof[2] = 0.2, 2.45
    p.set_state(dof, root)
    p.set_objective(task, np.asarray(goal, np.float32), None)
    return p


for name, p in (("pick 16 lanes", panda(64, 9, "pick")), ("pick 8 lanes", panda(44, 9, "pick", lanes=8)),
                ("reach mm 8 lanes + producer", panda(48, 9, "reach", lanes=8, mm=True)),
                ("pick thread per sample", panda(40, 9, "pick", lanes=1)),
                ("grasp state 8 lanes, halton table", panda(44, 12, "pick", lanes=8, grasp=True, noise=A.NOISE_TABLE)),
                ("grasp state 16 lanes", panda(30, 9, "pick", lanes=16, grasp=True)),
                ("grasp state thread per sample", panda(40, 9, "pick", lanes=1, grasp=True)),
                ("push_pull mm", point(64, 12, "push_pull", [-3.75, -3.75], mm=True))):
    a, c, info = p.command()
    assert np.isfinite(a).all() and np.isfinite(c).all(), name
    print("ok:", name, flush=True)
    p.close()
shards = [panda(64, 9, "pick", K_local=32, offset=32 * r) for r in range(2)]
desc = [s.peer_export() for s in shards]
for r, s in enumerate(shards):
    s.peer_attach(r, 2, desc)
for s in shards:
    s.command_resident()
for s in shards:
    a, c = s.fetch_result()
    assert np.isfinite(a).all()
print("ok: peer exchange, 2 handles", flush=True)
for s in shards:
    s.close()
