"""Generates tests/golden/halton_spline_*.npz with the UNMODIFIED reference sampler.

Run in the build container only (needs /root/reference and scipy):  python tests/golden/make_halton_golden.py

MPPI.get_samples (mppi.py:458-483) is called unbound on a namespace that carries exactly the attributes it reads; it
calls the reference's generate_gaussian_halton_samples (mppi_utils.py:99-104) and skill_utils.bspline
(skill_utils.py:9-22, scipy splrep / splev). The one missing piece is the un-vendored `ghalton` package
(pyproject.toml:15): the module the reference imports is a stand-in whose GeneralizedHalton restates the package's
published algorithm (digit-permuted radical inverse, dimension d in the d-th prime base, sequence index starting at 1).
ghalton.EA_PERMS itself is not obtainable offline, so the fixtures use two PINNED permutation sets that are stored in
the fixture: the identity (= the plain Halton sequence of the reference's own use_ghalton=False branch,
mppi_utils.py:82-87, which is cross-checked here) and a seeded random set with perm[0] = 0 like EA_PERMS.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "m3p2i-aip_b200"))
from reference_rig import import_reference  # noqa: E402

REF_SRC = "/root/reference/src"


def primes(n):
    out, c = [], 1
    while len(out) < n:
        c += 1
        if all(c % j for j in range(2, int(c ** 0.5) + 1)):
            out.append(c)
    return out


class GeneralizedHalton:
    """ghalton.GeneralizedHalton(perms).get(n): n points of the digit-scrambled Halton sequence, starting at index 1."""

    def __init__(self, perms):
        self.perms = [list(p) for p in perms]
        self.bases = primes(len(self.perms))
        self.count = 0

    def get(self, n):
        pts = []
        for _ in range(n):
            self.count += 1
            row = []
            for base, perm in zip(self.bases, self.perms):
                i, f, r = self.count, 1.0, 0.0
                while i > 0:
                    f /= base
                    r += f * perm[i % base]
                    i //= base
                row.append(r)
            pts.append(row)
        return pts


def pinned_perms(ndims, kind, stride):
    """uint16 [ndims, stride]: row d = permutation of 0 .. prime(d)-1 (padding zeros)."""
    out = np.zeros((ndims, stride), np.uint16)
    rng = np.random.default_rng(20240229)
    for d, b in enumerate(primes(ndims)):
        p = np.arange(b)
        if kind == "scrambled":
            p[1:] = rng.permutation(p[1:])   # perm[0] = 0, as in EA_PERMS
        out[d, :b] = p
    return out


def main():
    ref_m3p2i, _ = import_reference(REF_SRC)
    MPPI = ref_m3p2i.M3P2I.__mro__[1]   # the reference's MPPI class
    gh = types.ModuleType("ghalton")
    gh.GeneralizedHalton = GeneralizedHalton
    gen = MPPI.get_samples.__globals__["generate_gaussian_halton_samples"]
    gen.__globals__["ghalton"] = gh
    for T, K, nu in ((12, 48, 9), (16, 48, 9), (20, 64, 2), (32, 48, 9)):
        m = T // 4
        ndims = m * nu
        stride = primes(ndims)[-1]
        for kind in ("identity", "scrambled"):
            perms = pinned_perms(ndims, kind, stride)
            gh.EA_PERMS = [perms[d, :b].tolist() for d, b in enumerate(primes(ndims))]
            ns = types.SimpleNamespace(sampling_method="halton", ndims=ndims, seed_val=0, device="cpu", nu=nu, n_knots=m, T=T,
                                       degree=2, tensor_args={"device": "cpu", "dtype": torch.float32})
            delta = MPPI.get_samples(ns, K).numpy().astype(np.float32)
            knots = ns.knot_points.numpy().astype(np.float32)
            if kind == "identity":
                # the reference's own plain Halton branch gives the same knots (it accumulates the radical inverse in
                # fp32, ghalton in fp64: equal to a few fp32 ulp)
                plain = gen(K, ndims, use_ghalton=False, device="cpu", float_dtype=torch.float32).numpy()
                assert np.allclose(plain, knots, rtol=0, atol=1e-5), np.abs(plain - knots).max()
            np.savez_compressed(os.path.join(HERE, f"halton_spline_T{T}_{kind}.npz"), delta=delta, knots=knots, perms=perms,
                                K=K, T=T, nu=nu, knot_scale=4, degree=2, smoothing=0.5)
            print(f"T={T} K={K} nu={nu} {kind}: delta {delta.shape}, std {delta.std():.3f}")


if __name__ == "__main__":
    main()
