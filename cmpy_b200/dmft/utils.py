# -*- coding: utf-8 -*-
"""Host helpers of the DMFT loop (reference: cmpy/dmft/utils.py:66-280; the plotting helpers
of that file are out of scope)."""
import numpy as np

__all__ = ["IterationStats", "mix_values", "self_energy", "bethe_gf_omega", "quasiparticle_weight"]


class IterationStats:
    """Error history of a self-consistency loop + how it ended (reference: utils.py:66-160)."""

    def __init__(self, *names):
        self.names = list(names) or ["error"]
        self.errors = []
        self.status = ""
        self.success = False

    def append(self, *errs):
        self.errors.append(tuple(float(e) for e in errs))

    def __len__(self):
        return len(self.errors)

    def __getitem__(self, i):
        return self.errors[i]

    @property
    def num_iter(self):
        return len(self.errors)

    def set_parameter_converged(self, name, value):
        self.success = True
        self.status = f"{name} converged: {value}"

    def set_maxiter_status(self, max_iter):
        self.success = False
        self.status = f"maximum number of iterations reached ({max_iter})"

    def __str__(self):
        last = ", ".join(f"{n}={e:.2e}" for n, e in zip(self.names, self.errors[-1])) if self.errors else "-"
        return f"Iterations: {self.num_iter}\nStatus:     {self.status}\nLast error: {last}"


def mix_values(old, new, mixing=1.0):
    """Linear mixing; ``mixing == 1`` returns ``new`` unchanged (reference: utils.py:163-185)."""
    if mixing == 1:
        return new
    assert 0 < mixing < 1
    return new * mixing + old * (1.0 - mixing)


def self_energy(gf_imp0, gf_imp):
    """Sigma(z) = G0(z)^-1 - G(z)^-1 (reference: utils.py:188-211)."""
    return 1 / gf_imp0 - 1 / gf_imp


def bethe_gf_omega(z, t):
    """Local G of the infinitely coordinated Bethe lattice (reference: utils.py:214-230)."""
    z_rel = z / (2 * t)
    return z_rel * (1 - np.sqrt(1 - 1 / (z_rel * z_rel))) / t


def quasiparticle_weight(omegas, sigma, thresh=1e-5):
    """z_qp = 1 / (1 - dSigma/domega at 0): slope of Re Sigma fitted on the grid points with
    |omega| <= d_omega, zero below ``thresh`` (reference: utils.py:233-280)."""
    dw = omegas[1] - omegas[0]
    win = (-dw <= omegas) * (omegas <= +dw)
    try:
        dsigma = np.polyfit(omegas[win], np.real(sigma)[win], 1)[0]
    except np.linalg.LinAlgError:
        ow, sw = omegas[win], sigma[win]
        dsigma = (sw[-1] - sw[0]) / (ow[-1] - ow[0])
    z_qp = 1 / (1 - dsigma)
    if z_qp < thresh:
        z_qp = 0
    return z_qp
