"""Median device time of command_resident for C4 (panda pick K=4096 H=32) and C2/C3 (point), one line each."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200")); sys.path.insert(0, ROOT)
from m3p2i_b200 import _abi as A, native, scene as S
import bench

def run(env, task, K, T, mm=False, goal=None, robot=None, grip=None, q=None):
    cfg = S.make_cfg(env, task, goal, K, T, multi_modal=mm)
    if env == "panda_env":
        dof, root, g = bench.scene_inputs(); sc = S.build_panda_scene()
        if task != "pick": g = np.zeros(7, np.float32)
        if q is not None: dof = dof.copy(); dof[0::2] = q
    else:
        actors = S.default_actors(env); dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors)
        if robot: dof[0], dof[2] = robot
        g = np.asarray(goal, np.float32); sc = S.build_point_scene()
    p = native.NativePlanner(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), sc)
    p.set_filter_matrix(S.savgol_matrix(T)); p.set_state(dof, root); p.set_objective(task, g, grip)
    for _ in range(5): p.command_resident(sync=True)
    ms, rms = [], []
    for _ in range(30):
        i = p.command_resident(sync=True); ms.append(i.kernel_ms); rms.append(i.rollout_ms)
    p.close()
    print(f"{env} {task} K={K} T={T} mm={mm}: command {np.median(ms):.4f} ms, rollout {np.median(rms):.4f} ms, "
          f"{K*T/np.median(ms)/1e3:.1f} M sample-steps/s", flush=True)

if __name__ == "__main__":
    print("M3P2I_ROLLOUT_BLOCK =", os.environ.get("M3P2I_ROLLOUT_BLOCK"))
    run("panda_env", "pick", 4096, 32, grip="close")
    run("panda_env", "pick", 4096, 32, grip="close", q=[-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027])  # gripper closed around cubeA
    run("panda_env", "reach", 4096, 32, mm=True, grip="open")
    run("point_env", "navigation", 200, 12, goal=[-3.0, 3.0])
    run("point_env", "push", 1024, 20, goal=[-1.0, -1.0], robot=[0.2, 2.45])
    run("point_env", "push_pull", 4096, 20, mm=True, goal=[-3.75, -3.75], robot=[0.3, 2.4])
