"""K-sweep of the device-resident command (Panda pick / point push, Philox noise): where does throughput saturate?
Prints one CSV line per K: env,K,H,ms_per_command,rollout_ms,Msample_steps_per_s,rollout_GBps(algorithmic)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200"))
sys.path.insert(0, ROOT)

from m3p2i_b200 import _abi as A  # noqa: E402
from m3p2i_b200 import native  # noqa: E402
from m3p2i_b200 import scene as S  # noqa: E402
import bench  # noqa: E402


def run(env, K, T, reps=20):
    if env == "panda_env":
        cfg = S.make_cfg(env, "pick", None, K, T)
        dof, root, goal = bench.scene_inputs()
        task, grip, sc, nbytes = "pick", "close", S.build_panda_scene(), 56
    else:
        cfg = S.make_cfg(env, "push", [-1.0, -1.0], K, T)
        actors = S.default_actors(env)
        dof, root = S.initial_dof_state(actors).copy(), S.initial_root_state(actors)
        dof[0], dof[2] = 0.2, 2.45
        goal, task, grip, sc, nbytes = np.array([-1.0, -1.0], np.float32), "push", None, S.build_point_scene(), 28
    p = native.NativePlanner(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), sc)
    p.set_filter_matrix(S.savgol_matrix(T))
    p.set_state(dof, root)
    p.set_objective(task, goal, grip)
    for _ in range(3):
        p.command_resident(sync=True)
    ms, rms = [], []
    for _ in range(reps):
        info = p.command_resident(sync=True)
        ms.append(info.kernel_ms)
        rms.append(info.rollout_ms)
    p.close()
    m, r = float(np.median(ms)), float(np.median(rms))
    print(f"{env},{K},{T},{m:.4f},{r:.4f},{K * T / m / 1e3:.1f},{nbytes * K * T / r / 1e6:.1f}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1:
    print("env,K,H,ms_per_command,rollout_ms,Msample_steps_per_s,rollout_algorithmic_GBps  (M3P2I_LANES=%s)" % os.environ.get("M3P2I_LANES"))
    for K in [int(a) for a in sys.argv[1:]]:
        run("panda_env", K, 32)
    sys.exit(0)

if __name__ == "__main__":
    print("env,K,H,ms_per_command,rollout_ms,Msample_steps_per_s,rollout_algorithmic_GBps")
    for K in (1024, 4096, 16384, 65536, 262144, 1048576):
        run("panda_env", K, 32)
    for K in (1024, 4096, 16384, 65536, 262144, 1048576, 4194304):
        run("point_env", K, 20)
