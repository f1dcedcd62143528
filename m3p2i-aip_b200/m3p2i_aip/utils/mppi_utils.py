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


# ---------------------------------------------------------------------------------------------------------------------
# Small public helpers of the reference module that user code may import. Inside command() both run in the CUDA
# kernels (clamping in the rollout kernel, discounting in its cost accumulation); these host versions exist for callers.
def scale_ctrl(ctrl, action_lows, action_highs, squash_fn="clamp"):
    """Bound a control sequence (mppi_utils.py:28-44): "clamp" to [lows, highs]; "clamp_rescale" / "tanh" map [-1, 1] onto
    the range; "identity" returns the input. 1-D input is treated as [1, nu, 1] like the reference does."""
    import torch
    if ctrl.dim() == 1:
        ctrl = ctrl.view(1, -1, 1)
    if squash_fn == "identity":
        return ctrl
    if squash_fn == "clamp":
        return torch.maximum(torch.minimum(ctrl, action_highs), action_lows)
    if squash_fn == "clamp_rescale":
        unit = ctrl.clamp(-1.0, 1.0)
    elif squash_fn == "tanh":
        unit = torch.tanh(ctrl)
    else:
        raise ValueError(f"unknown squash_fn {squash_fn!r}")
    mid, half = 0.5 * (action_highs + action_lows), 0.5 * (action_highs - action_lows)
    return mid.unsqueeze(0) + unit * half.unsqueeze(0)


def cost_to_go(cost_seq, gamma_seq):
    """Discounted cost-to-go of every step (mppi_utils.py:106-113): out[:, t] = sum_{s >= t} gamma^(s - t) * c[:, s]
    for gamma_seq = gamma^[0..T-1]; column 0 is the discounted trajectory cost the softmin uses."""
    import torch
    weighted = gamma_seq * cost_seq
    tail_sums = torch.flip(torch.cumsum(torch.flip(weighted, dims=[-1]), dim=-1), dims=[-1])
    return tail_sums / gamma_seq
