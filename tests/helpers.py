"""Shared helpers of the parity tests: replay a golden case on any backend with the planner method surface."""
import glob
import os

import numpy as np

from m3p2i_b200 import _abi as A
from m3p2i_b200 import scene as S

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRIPPER = {"reach": "open", "place": "open", "pick": "close"}  # m3p2i.py:10-14


def golden_cases():
    # planner cases only; the halton_spline_* fixtures of the noise table have their own test (test_halton_spline.py)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith("halton_spline_"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def case_cfg(g, K=None, T=None):
    env = str(g["env"])
    goal = g["goal"].tolist()
    return S.make_cfg(env, str(g["task"]), goal, int(K or g["K"]), int(T or g["T"]), multi_modal=bool(g["multi_modal"]),
                      cube_on_shelf=bool(g["cube_on_shelf"]), device="cpu")


def build_scene(env, actors=None):
    return S.build_point_scene(actors) if env == "point_env" else S.build_panda_scene(actors)


def make_backend(cls, cfg, noise_mode=A.NOISE_TABLE, seed=0, K_local=None, offset=0, **kw):
    c = S.build_config(cfg, num_samples_local=K_local, sample_offset=offset, noise_mode=noise_mode, seed=seed)
    b = cls(c, build_scene(cfg.env_type), **kw)
    if cfg.mppi.filter_u:
        b.set_filter_matrix(S.savgol_matrix(int(cfg.mppi.horizon)))
    return b


def tick(b, g, i):
    """One run_tamp tick of golden case g, call i, on backend b (reactive_tamp.py:43-61)."""
    task = str(g["task"])
    b.set_state(g[f"dof_{i}"], g[f"root_{i}"])
    if bool(g["extra_step"]):
        b.sim_step()
    b.set_objective(task, g["goal"], GRIPPER.get(task) if str(g["env"]) == "panda_env" else None)
    return b.command()


def assert_close(a, b, rtol, atol, what, max_bad_frac=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = ~np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    frac = bad.mean() if bad.size else 0.0
    if frac > max_bad_frac:
        idx = np.argwhere(bad)[:5]
        detail = ", ".join(f"{tuple(i)}: {a[tuple(i)]:.6g} vs {b[tuple(i)]:.6g}" for i in idx)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} elements differ (rtol={rtol}, atol={atol}); {detail}")
