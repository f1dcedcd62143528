"""The halton-spline noise table (MPPI.get_samples, mppi.py:458-478): the C restatement (oracle/halton_spline.h) and the
device builder (m3p2i_set_noise_halton_spline, csrc/halton_spline.cuh) against fixtures written by the unmodified
reference sampler (tests/golden/make_halton_golden.py) and against scipy's splrep / splev on random knot vectors."""
import glob
import os
import warnings

import numpy as np
import pytest

import oracle_py as O
from helpers import make_backend
from m3p2i_b200 import _abi as A
from m3p2i_b200 import scene as S

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "halton_spline_T*.npz")))
TOL = 1e-5   # absolute, on unit-variance noise


def _load(path):
    g = np.load(path)
    return g, int(g["K"]), int(g["T"]), int(g["nu"]), int(g["knot_scale"]), int(g["degree"]), float(g["smoothing"])


def test_fixtures_exist():
    assert len(GOLDEN) == 8   # T in {12, 16, 20, 32} x {identity, scrambled} permutations


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_table_matches_reference_sampler(path):
    g, K, T, nu, ks, deg, s = _load(path)
    perms = None if path.endswith("identity.npz") else g["perms"]
    d = O.halton_spline_table(K, T, nu, ks, deg, s, perms=perms)
    assert np.abs(d - g["delta"]).max() < TOL
    # identity permutations passed explicitly are the plain sequence
    if perms is None:
        assert np.array_equal(d, O.halton_spline_table(K, T, nu, ks, deg, s, perms=g["perms"]))
    # a shard (global samples 16 .. 32) is a slice of the table
    assert np.array_equal(O.halton_spline_table(16, T, nu, ks, deg, s, offset=16, perms=perms), d[16:32])


@pytest.mark.parametrize("m,T", [(3, 12), (4, 16), (5, 20), (8, 32), (7, 30), (16, 64)])
def test_oracle_spline_matches_scipy(m, T):
    """skill_utils.bspline = splrep(linspace(0, m, m), cv, k=2, s=0.5) + splev(linspace(0, m, T), ext=3): every branch of
    FITPACK's curfit (polynomial accepted, knots added one / several at a time, interpolation knots, the root search
    for the smoothing parameter) on random data of different scales."""
    import scipy.interpolate as si
    rng = np.random.default_rng(m)
    worst = 0.0
    for _ in range(400):
        cv = (rng.normal(size=m) * rng.choice([0.1, 1.0, 3.0])).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spl = si.splrep(np.linspace(0, m, m), cv, k=2, s=0.5)
        ref = si.splev(np.linspace(0, m, T), spl, ext=3).astype(np.float32)
        worst = max(worst, float(np.abs(O.bspline_samples(cv, T) - ref).max()))
    assert worst < 1e-6, worst


def test_host_mirror_table_is_the_same_table():
    """m3p2i_aip.utils.mppi_utils.halton_spline_table (host, scipy) and the C restatement agree."""
    from m3p2i_aip.utils import mppi_utils
    assert np.abs(mppi_utils.halton_spline_table(32, 16, 9) - O.halton_spline_table(32, 16, 9)).max() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_device_table_matches_reference_sampler(path):
    from m3p2i_b200 import native
    g, K, T, nu, ks, deg, s = _load(path)
    env = "panda_env" if nu == 9 else "point_env"
    cfg = S.make_cfg(env, "reach" if nu == 9 else "navigation", None if nu == 9 else [1.0, 1.0], K, T)
    perms = None if path.endswith("identity.npz") else g["perms"]
    n = make_backend(native.NativePlanner, cfg)
    n.set_noise_halton_spline(ks, deg, s, perms=perms)
    d = n.get_noise()
    n.close()
    assert np.abs(d - g["delta"]).max() < TOL
    # the same fp64 algorithm on both sides (the device contracts multiply-adds: equal to rounding of the fp32 result)
    assert np.abs(d - O.halton_spline_table(K, T, nu, ks, deg, s, perms=perms)).max() < 1e-6
    # a shard builds its own rows (global sample ids) and the row of global sample 0
    sh = make_backend(native.NativePlanner, cfg, K_local=16, offset=16)
    sh.set_noise_halton_spline(ks, deg, s, perms=perms)
    assert np.array_equal(sh.get_noise(), d[16:32])
    sh.close()


@pytest.mark.gpu
def test_device_table_at_planner_size_is_fast_and_sane():
    """C4 size: 4096 x 9 splines (the reference: 36 864 scipy calls) in one launch; unit-ish variance, smooth in time."""
    import time
    from m3p2i_b200 import native
    cfg = S.make_cfg("panda_env", "pick", None, 4096, 32)
    n = make_backend(native.NativePlanner, cfg)
    n.set_noise_halton_spline()
    t0 = time.perf_counter()
    n.set_noise_halton_spline()
    dt = time.perf_counter() - t0
    d = n.get_noise()
    n.close()
    assert dt < 0.25, dt
    assert 0.6 < d.std() < 1.1
    assert np.abs(np.diff(d, axis=1)).mean() < 0.25 * d.std()
    ref = O.halton_spline_table(256, 32, 9, offset=2048)
    assert np.abs(d[2048:2304] - ref).max() < 1e-6


@pytest.mark.gpu
def test_device_rejects_bad_arguments():
    from m3p2i_b200 import native
    cfg = S.make_cfg("panda_env", "pick", None, 64, 12)
    n = make_backend(native.NativePlanner, cfg)
    with pytest.raises(native.NativeError):
        n.set_noise_halton_spline(degree=3)                                # 12 // 4 = 3 knot points <= degree
    with pytest.raises(native.NativeError):
        n.set_noise_halton_spline(perms=np.zeros((27, 103), np.uint16))   # rows are not permutations
    with pytest.raises(native.NativeError):
        n.set_noise_halton_spline(perms=np.zeros((27, 50), np.uint16))    # stride below the largest base (103)
    n.set_noise_halton_spline()
    n.close()
    ph = make_backend(native.NativePlanner, cfg, noise_mode=A.NOISE_PHILOX)
    with pytest.raises(native.NativeError):
        ph.set_noise_halton_spline()                                       # in-kernel sampling: a table would be ignored
    ph.close()
