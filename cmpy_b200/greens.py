# -*- coding: utf-8 -*-
"""Green's-function helper on the hot path (reference: cmpy/greens.py:18-64).

Only ``gf0_lehmann`` -- the non-interacting Lehmann sum used as the U=0 oracle of G(z) --
is in scope; the closed-form lattice Green's functions of the reference module are not
(SURVEY.md section 2, #11).  The pole sum runs in the ``cmpy_pole_sum`` kernel."""
from typing import Union

import numpy as np

from . import _lib
from .exactdiag import pole_sum, _z_tensor

__all__ = ["gf0_lehmann"]


def gf0_lehmann(*args, z: Union[complex, np.ndarray], mu: float = 0.0, mode="diag") -> np.ndarray:
    """Non-interacting Green's function from a Hamiltonian matrix (one argument) or from
    ``(eigvals, eigvecs)``; ``mode`` in ``'full' | 'diag' | 'total'``.  The index contraction
    follows the reference's einsum strings verbatim (cmpy/greens.py:51-64)."""
    if len(args) == 1:
        eigvals, eigvecs = np.linalg.eigh(args[0])  # N x N single-particle problem (host)
    else:
        eigvals, eigvecs = args
    if mode not in ("full", "diag", "total"):
        raise ValueError(f"Mode '{mode}' not supported. Valid modes are 'full', 'diag' or 'total'")
    if np.iscomplexobj(eigvecs):
        raise NotImplementedError("complex eigenvectors are not supported by the device pole sum")
    torch = _lib.require_cuda()
    dev = _lib.device()
    z = np.atleast_1d(z)
    zt, zshape = _z_tensor(z)
    vecs = np.asarray(eigvecs, dtype=np.float64)
    poles = torch.from_numpy(np.ascontiguousarray(np.asarray(eigvals, dtype=np.float64) - mu)).to(dev)
    n = vecs.shape[0]

    def psum(weights):
        w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
        return pole_sum(w, poles, zt).cpu().numpy().reshape(zshape)

    if mode == "diag":      # out[..., i] = sum_j vecs[j, i]^2 / (z + mu - eps_j)
        return np.stack([psum(vecs[:, i] ** 2) for i in range(n)], axis=-1)
    if mode == "total":     # sum over i of the above
        return psum((vecs ** 2).sum(axis=1))
    out = np.empty(zshape + (n, n), dtype=np.complex128)   # 'full': sum_k v[k,i] v[k,j] / arg_k
    for i in range(n):
        for j in range(n):
            out[..., i, j] = psum(vecs[:, i] * vecs[:, j])
    return out
