# -*- coding: utf-8 -*-
"""Two-site DMFT self-consistency at half filling (reference: cmpy/dmft/twosite.py:25-178).

The impurity problem is the two-site SIAM; ``ref=True`` uses the closed-form T=0 poles and
residues (E. Lange), ``ref=False`` the exact-diagonalisation solver: the finite-temperature
Lehmann sum ``siam.impurity_gf`` (GPU eigensolver + pole-sum kernel) for finite ``beta`` and
the GPU Lanczos continued fraction for ``beta = inf`` (where the reference divides by zero)."""
import numpy as np

from ..models.anderson import SingleImpurityAndersonModel
from .utils import IterationStats, self_energy, quasiparticle_weight, mix_values, bethe_gf_omega


def impurity_params_ref(u, v):
    """Residues a1, a2 and poles e1, e2 of the two-site SIAM G(z) at half filling, T=0
    (reference: twosite.py:25-60)."""
    r16 = np.sqrt(u ** 2 + 16 * v ** 2)
    r64 = np.sqrt(u ** 2 + 64 * v ** 2)
    a1 = 0.25 * (1 - (u ** 2 - 32 * v ** 2) / np.sqrt((u ** 2 + 64 * v ** 2) * (u ** 2 + 16 * v ** 2)))
    a2 = 0.5 - a1
    return float(a1), float(a2), float(0.25 * (r64 - r16)), float(0.25 * (r64 + r16))


def impurity_gf_ref(z, u, v):
    """Closed-form impurity G(z) of the two-site SIAM (reference: twosite.py:63-82)."""
    a1, a2, e1, e2 = impurity_params_ref(u, v)
    return (a1 / (z - e1) + a1 / (z + e1)) + (a2 / (z - e2) + a2 / (z + e2))


def impurity_gf0(z, siam):
    return siam.impurity_gf0(z)


def compute_impurity_gf(z, siam, ref=False):
    if ref:
        return impurity_gf_ref(z, siam.u, siam.v)
    if siam.temp == 0:
        from ..exactdiag import gf_continued_fraction

        n = siam.num_sites // 2
        return gf_continued_fraction(siam, z, pos=0, n_up=n, n_dn=siam.num_sites - n)
    return siam.impurity_gf(z)


def compute_self_energy(z, siam, ref=False):
    return self_energy(impurity_gf0(z, siam), compute_impurity_gf(z, siam, ref=ref))


def twosite_dmft_half_filling(z, u, t=1.0, beta=np.inf, mixing=1.0, vtol=1e-6, max_iter=1000,
                              vthresh=1e-10, verbose=True, ref=True):
    """Iterates V -> sqrt(z_qp) t until |dV| < vtol; returns the converged SIAM
    (reference: twosite.py:103-172)."""
    siam = SingleImpurityAndersonModel(u, v=[t], mu=u / 2, temp=1 / beta)
    v = siam.v[0] + 0.1  # must differ from the current value, or the first error is zero
    m2 = t ** 2
    it = 0
    stats = IterationStats("Δv")
    while True:
        siam.update_hybridization(v)
        sigma = self_energy(impurity_gf0(z, siam), compute_impurity_gf(z, siam, ref=ref))
        qp_weight = quasiparticle_weight(z.real, sigma, thresh=vthresh)
        v_new = mix_values(v, np.sqrt(qp_weight * m2), mixing=mixing)
        delta_v = np.linalg.norm(v - v_new)
        stats.append(delta_v)
        v = v_new
        if v == 0:
            stats.set_parameter_converged("Hybridization", v)
            break
        if delta_v < vtol:
            stats.set_parameter_converged("Hybridization", v)
            break
        elif qp_weight == 0:
            stats.set_parameter_converged("Quasiparticle weight", qp_weight)
            break
        elif it >= max_iter:
            stats.set_maxiter_status(max_iter)
            break
        it += 1
    siam.update_hybridization(v)
    if verbose:
        print("-" * 50)
        print(f"U:          {u:.2f}")
        print(stats)
    return siam


def compute_lattice_greens_function(z, siam, t, ref=False):
    sigma = compute_self_energy(z, siam, ref)
    return bethe_gf_omega(z + siam.mu - sigma, t)
