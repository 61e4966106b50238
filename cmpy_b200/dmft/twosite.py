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


def _next_hybridization(z, siam, v, t, mixing, vthresh, ref):
    """One self-consistency step: solve the impurity problem at hybridisation ``v``, return the mixed
    new hybridisation sqrt(Z) t and the quasiparticle weight Z (reference: twosite.py:127-139)."""
    siam.update_hybridization(v)
    sigma = compute_self_energy(z, siam, ref=ref)
    weight = quasiparticle_weight(z.real, sigma, thresh=vthresh)
    return mix_values(v, np.sqrt(weight * t * t), mixing=mixing), weight


def twosite_dmft_half_filling(z, u, t=1.0, beta=np.inf, mixing=1.0, vtol=1e-6, max_iter=1000,
                              vthresh=1e-10, verbose=True, ref=True):
    """Two-site DMFT of the half-filled Hubbard model on the Bethe lattice: the hybridisation of the
    two-site SIAM is iterated, V -> sqrt(Z) t, until it moves by less than ``vtol`` (or vanishes, or the
    quasiparticle weight does, or ``max_iter`` is hit); returns the converged SIAM.  Same stopping rules,
    order of checks and messages as the reference loop (twosite.py:103-172)."""
    siam = SingleImpurityAndersonModel(u, v=[t], mu=u / 2, temp=1 / beta)
    stats = IterationStats("Δv")
    v = siam.v[0] + 0.1   # start away from the model's own value: the first step must register a change
    for it in range(max_iter + 1):
        v_old = v
        v, weight = _next_hybridization(z, siam, v_old, t, mixing, vthresh, ref)
        change = np.linalg.norm(v_old - v)
        stats.append(change)
        if v == 0 or change < vtol:
            stats.set_parameter_converged("Hybridization", v)
            break
        if weight == 0:
            stats.set_parameter_converged("Quasiparticle weight", weight)
            break
        if it >= max_iter:
            stats.set_maxiter_status(max_iter)
    siam.update_hybridization(v)
    if verbose:
        print("-" * 50 + f"\nU:          {u:.2f}\n{stats}")
    return siam


def compute_lattice_greens_function(z, siam, t, ref=False):
    sigma = compute_self_energy(z, siam, ref)
    return bethe_gf_omega(z + siam.mu - sigma, t)
