# -*- coding: utf-8 -*-
"""Spin observables of a state on a magnetisation sector (SURVEY.md section 8(f), row f-4;
reference: the helper functions of scripts/heisenberg.py:50-143).

``<Sz_i Sz_j>`` is the expectation value of a diagonal operator that the Heisenberg kernel
already knows: a model with the single directed pair (i, j), ``j = 0`` and ``jz = 1`` has the
diagonal ``(-1)^(b_i + b_j) / 4 = Sz_i Sz_j`` (cmpy/models/heisenberg.py:28-31), so the correlator
is ``sum_s psi_s^2 diag_s`` with the diagonal produced on the device by ``cmpy_op_diagonal``
(kernel K5, works up to 32 sites without materialising the state list).  ``<Sz_i>`` is a
one-line reduction over the enumerated states (kernel K1).  No CPU path: the functions need a
CUDA device like the rest of the package."""
import math

import numpy as np

from . import _lib
from .basis import enumerate_states
from .operators import SpinHamiltonOperator

__all__ = ["sz_diagonal", "sz_sz_diagonal", "sz_expval", "sz_correl", "spin_correlations"]


def _as_device_vector(gs):
    torch = _lib.require_cuda()
    if isinstance(gs, torch.Tensor):
        return gs.to(device=_lib.device(), dtype=torch.float64).reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(gs, dtype=np.float64).reshape(-1))).to(_lib.device())


def _n_up(num_sites, s):
    n_up = num_sites / 2 + s
    if n_up % 1 != 0.0 or not (0 <= n_up <= num_sites):
        raise ValueError(f"Total spin of {s} not realizable with {num_sites} sites")  # cmpy/basis.py:755-758
    return int(n_up)


def sz_sz_diagonal(num_sites, s, i, j):
    """Device tensor ``Sz_i Sz_j`` (= +-1/4) for every state of the sector with total spin ``s``."""
    torch = _lib.require_cuda()
    if i == j:
        n = math.comb(num_sites, _n_up(num_sites, s))
        return torch.full((n,), 0.25, dtype=torch.float64, device=_lib.device())
    op = SpinHamiltonOperator(num_sites, _n_up(num_sites, s), [(int(i), int(j))], 0.0, 1.0)
    d = torch.empty(op.shape[0], dtype=torch.float64, device=_lib.device())
    _lib.check(_lib.lib().cmpy_op_diagonal(op.handle, _lib.ptr(d), _lib.stream_ptr()), "cmpy_op_diagonal")
    return d


def sz_diagonal(num_sites, s, pos):
    """Device tensor ``Sz_pos`` (= +-1/2) for every state of the sector (state list from K1)."""
    torch = _lib.require_cuda()
    states = torch.from_numpy(np.asarray(enumerate_states(num_sites, _n_up(num_sites, s)), dtype=np.int64))
    states = states.to(_lib.device())
    return ((states >> int(pos)) & 1).to(torch.float64) - 0.5


def sz_expval(num_sites, s, gs, pos=0):
    """``<gs| Sz_pos |gs>`` (reference: ``sz_expval``, scripts/heisenberg.py:50-57: sign +1/2 for a
    set bit, -1/2 otherwise, weighted with the squared amplitudes)."""
    psi = _as_device_vector(gs)
    return float((psi * psi * sz_diagonal(num_sites, s, pos)).sum())


def sz_correl(num_sites, s, gs, delta, j=1.0, pos=0):
    """``j <gs| Sz_pos Sz_(pos+delta) |gs>`` (reference: ``sz_correl``, scripts/heisenberg.py:131-138,
    which fixes ``pos = 0``: sum of ``a^2 * sign * j / 4``, sign +1 for equal bits)."""
    psi = _as_device_vector(gs)
    return float(j) * float((psi * psi * sz_sz_diagonal(num_sites, s, pos, pos + delta)).sum())


def spin_correlations(num_sites, s, gs, pos=0):
    """``[<Sz_pos Sz_k> for k in range(num_sites)]`` as a numpy array."""
    psi = _as_device_vector(gs)
    w = psi * psi
    return np.asarray([float((w * sz_sz_diagonal(num_sites, s, pos, k)).sum()) for k in range(num_sites)])
