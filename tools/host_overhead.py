"""Where the host-side time of one e2e command goes (C4): enqueue only, enqueue + sync, set_state, full command."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200")); sys.path.insert(0, ROOT)
from m3p2i_b200 import _abi as A, native, scene as S
import bench
cfg = S.make_cfg("panda_env", "pick", None, 4096, 32)
dof, root, g = bench.scene_inputs()
p = native.NativePlanner(S.build_config(cfg, noise_mode=A.NOISE_PHILOX, seed=0), S.build_panda_scene())
p.set_filter_matrix(S.savgol_matrix(32)); p.set_state(dof, root); p.set_objective("pick", g, "close")
for _ in range(20): p.command()
def t(fn, n=300):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
print(f"set_state                 {t(lambda: p.set_state(dof, root)):8.1f} us")
print(f"command (sync, D2H)       {t(lambda: p.command()):8.1f} us")
print(f"command(want_cost=False)  {t(lambda: p.command(want_cost=False)):8.1f} us")
def both():
    p.set_state(dof, root); p.command(want_cost=False)
print(f"set_state + command       {t(both):8.1f} us")
i = p.command_resident(sync=True)
print(f"device time of a command  {1e3 * i.kernel_ms:8.1f} us (rollout {1e3 * i.rollout_ms:.1f})")
fn = p.fn["m3p2i_version"]
print(f"empty ctypes call         {t(lambda: fn(), 2000):8.2f} us")
