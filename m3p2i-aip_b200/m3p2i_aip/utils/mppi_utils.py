"""Host-side sampling helpers of the planner (reference: utils/mppi_utils.py).

Only the one-time noise-table construction lives here; clamping, discounting (cost_to_go) and the softmin run in
the CUDA kernels. The reference draws its knots from the `ghalton` package's generalised Halton sequence with
Evolutionary-Algorithm permutations (mppi_utils.py:88-95); that package is a third-party C++/SWIG dependency that
is not vendored, so the default table here uses the plain (radical-inverse) Halton sequence that the reference
implements itself for use_ghalton=False (mppi_utils.py:68-87). Same marginals, different low-discrepancy points.
"""
import numpy as np


def generate_prime_numbers(num):
    primes, n = [], 2
    while len(primes) < num:
        if all(n % p for p in primes if p * p <= n):
            primes.append(n)
        n += 1 if n == 2 else 2
    return primes


def generate_halton_samples(num_samples, ndims, bases=None):
    """[num_samples, ndims] radical-inverse Halton points, index starting at 1 (mppi_utils.py:68-87)."""
    bases = bases or generate_prime_numbers(ndims)
    out = np.zeros((num_samples, ndims), np.float64)
    idx0 = np.arange(1, num_samples + 1, dtype=np.int64)
    for d in range(ndims):
        base, f, idx, r = bases[d], 1.0, idx0.copy(), np.zeros(num_samples)
        while (idx > 0).any():
            f /= base
            r += f * (idx % base)
            idx //= base
        out[:, d] = r
    return out


def generate_gaussian_halton_samples(num_samples, ndims, bases=None):
    """sqrt(2) * erfinv(2u - 1) of the Halton points (mppi_utils.py:99-104)."""
    from scipy.special import erfinv
    u = generate_halton_samples(num_samples, ndims, bases)
    return (np.sqrt(2.0) * erfinv(2.0 * u - 1.0)).astype(np.float32)


def bspline(c_arr, n=100, degree=3):
    """Smoothing spline through the knots, resampled at n points (skill_utils.py:9-22: splrep(k=degree, s=0.5))."""
    import scipy.interpolate as si
    cv = np.asarray(c_arr, np.float64)
    t_arr = np.linspace(0, cv.shape[0], cv.shape[0])
    spl = si.splrep(t_arr, cv, k=degree, s=0.5)
    return si.splev(np.linspace(0, cv.shape[0], n), spl, ext=3)


def halton_spline_table(K, T, nu, knot_scale=4, degree=2):
    """The once-sampled noise table delta [K,T,nu] of the halton-spline mode (mppi.py:458-478)."""
    n_knots = T // knot_scale
    if n_knots <= degree:
        raise ValueError(f"horizon {T} gives {n_knots} knots; the degree-{degree} spline needs more "
                         "(the reference YAMLs say: at least 12)")
    knots = generate_gaussian_halton_samples(K, n_knots * nu).reshape(K, nu, n_knots)
    out = np.zeros((K, T, nu), np.float32)
    for i in range(K):
        for j in range(nu):
            out[i, :, j] = bspline(knots[i, j], n=T, degree=degree)
    return out
