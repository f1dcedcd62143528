"""Grasp-state parity of every rollout kernel shape against the oracle (GPU): fraction of samples whose step costs
differ, per tick, with the planner state of the oracle copied into the CUDA planner before each tick (sync=1: same
inputs every tick) or left to evolve on its own (sync=0).   python tests/experiments/grasp_parity.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT]
import oracle_py as O
from helpers import make_backend
from m3p2i_b200 import _abi as A, native, scene as S
O.set_threads(os.cpu_count())
GRASP_Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027]
K, T = 512, 16
for sync in (1, 0):
    for lanes in (1, 8, 16):
        cfg = S.make_cfg("panda_env", "pick", None, K, T)
        cfg.mppi.lanes_per_sample = lanes
        o = make_backend(O.Oracle, cfg, noise_mode=A.NOISE_PHILOX, seed=7)
        n = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX, seed=7)
        actors = S.default_actors("panda_env")
        dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors).copy()
        dof[0::2] = GRASP_Q
        root[S.actor_index(actors, "cubeA"), 2] -= 0.0095; root[S.actor_index(actors, "cubeB"), 2] -= 0.0095
        cb = root[S.actor_index(actors, "cubeB")]
        goal = np.concatenate([cb[:3] + np.array([0, 0, 0.055], np.float32), cb[3:7]])
        for b in (o, n):
            b.set_state(dof, root); b.set_objective("pick", goal, "close")
        for i in range(3):
            if sync: n.set_planner_state(o.get_planner_state())
            a_n, _, _ = n.command(); a_o, _, _ = o.command()
            ch_n, ch_o = n.read_buffer(A.BUF_COST_HORIZON), o.read_buffer(A.BUF_COST_HORIZON)
            d = np.abs(ch_n - ch_o)
            bad = ~np.isclose(ch_n, ch_o, rtol=1e-3, atol=1e-3)
            flips = ((ch_n > 900) != (ch_o > 900))
            print(f"sync={sync} lanes={lanes} tick {i}: samples differing {bad.any(1).mean():.4f} (collision flips in {flips.any(1).mean():.4f}), "
                  f"median |dc| {np.median(d):.2e}, 99% {np.quantile(d, 0.99):.2e}, action max diff {np.abs(a_n - a_o).max():.2e}, collision steps {int((ch_o > 900).sum())}/{ch_o.size}", flush=True)
        o.close(); n.close()
