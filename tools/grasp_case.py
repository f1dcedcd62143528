"""One device-resident command of the contact-rich Panda pick state (fingers closed around cubeA) for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200")); sys.path.insert(0, ROOT)
from m3p2i_b200 import _abi as A, native, scene as S
import bench
Q = [-0.21448, 1.040633, -0.091726, -1.527796, 0.144743, 2.56178, 0.807333, 0.027, 0.027]
cfg = S.make_cfg("panda_env", "pick", None, 4096, 32)
dof, root, g = bench.scene_inputs(); dof = dof.copy(); dof[0::2] = Q
p = native.NativePlanner(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), S.build_panda_scene())
p.set_filter_matrix(S.savgol_matrix(32)); p.set_state(dof, root); p.set_objective("pick", g, "close")
for _ in range(4):
    i = p.command_resident(sync=True)
print("rollout_ms", i.rollout_ms)
st = p.read_buffer(A.BUF_STATES); ch = p.read_buffer(A.BUF_COST_HORIZON)
print("cost_h mean/min/max", ch.mean(), ch.min(), ch.max(), "nan", np.isnan(ch).sum())
