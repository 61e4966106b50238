# -*- coding: utf-8 -*-
"""Hubbard model (reference: cmpy/models/hubbard.py).

H = U sum_i n_iup n_idn + sum_{i,sigma} (eps_i - mu) n_{i sigma} + t sum_{<ij>,sigma} c+_i c_j
with the reference's conventions: matrix element ``+hop * sign`` (cmpy/operators.py:454),
only neighbor pairs with ``i < j`` are used (cmpy/models/hubbard.py:20-22), on-site energy
``eps - mu`` (hubbard.py:77)."""
import numpy as np
from scipy.sparse import csr_matrix

from ..operators import project_hubbard_inter, project_onsite_energy, project_hopping
from .abc import AbstractManyBodyModel

__all__ = ["HubbardModel", "hubbard_hamiltonian"]


def _ham_data(up_states, dn_states, num_sites, neighbors, inter, eps, hop):
    """COO triplets of the Hubbard Hamiltonian in the reference's emission order: on-site
    energies, interaction, then one hopping block per neighbor pair (i < j)."""
    yield from project_onsite_energy(up_states, dn_states, np.full(num_sites, eps))
    yield from project_hubbard_inter(up_states, dn_states, np.full(num_sites, inter))
    for i, j in neighbors:
        if i < j:
            yield from project_hopping(up_states, dn_states, num_sites, i, j, hop)


def hubbard_hamiltonian(sector, neighbors, inter=0.0, eps=0.0, hop=1.0):
    """The reference's scipy path: Hamiltonian of one sector as ``csr_matrix``
    (reference: cmpy/models/hubbard.py:25-34)."""
    rows, cols, data = list(), list(), list()
    for i, j, val in _ham_data(sector.up_states, sector.dn_states, sector.num_sites, neighbors,
                               inter, eps, hop):
        rows.append(i)
        cols.append(j)
        data.append(val)
    return csr_matrix((data, (rows, cols)))


class HubbardModel(AbstractManyBodyModel):
    """``HubbardModel(latt, ...)`` or ``HubbardModel(num_sites, neighbors, ...)`` with
    ``inter`` (U), ``eps``, ``hop`` and ``mu`` (reference: cmpy/models/hubbard.py:37-81)."""

    def __init__(self, *args, inter=0.0, eps=0.0, hop=1.0, mu=0.0):
        if len(args) == 1:
            latt = args[0]
            num_sites = latt.num_sites
            neighbors = latt.neighbor_pairs(True)[0]
        else:
            num_sites, neighbors = args
        super().__init__(num_sites, inter=inter, eps=eps, hop=hop, mu=mu)
        self.neighbors = neighbors

    def pformat(self):
        return f"U={self.inter}, ε={self.eps}, t={self.hop}, μ={self.mu}"

    def _hamiltonian_data(self, up_states, dn_states):
        return _ham_data(up_states, dn_states, self.num_sites, self.neighbors, self.inter,
                         self.eps - self.mu, self.hop)

    def _operator_spec(self):
        n = self.num_sites
        bonds = [(int(i), int(j)) for i, j in self.neighbors if i < j]
        return dict(bonds=bonds, hops=np.full(len(bonds), self.hop, dtype=np.float64),
                    eps=np.full(n, self.eps - self.mu, dtype=np.float64),
                    u=np.full(n, self.inter, dtype=np.float64), sign_width=n)
