# -*- coding: utf-8 -*-
"""Heisenberg / XXZ model (reference: cmpy/models/heisenberg.py).

Reference scaling (pinned by cmpy/tests/test_models_heisenberg.py:15-35): every *directed*
neighbor pair contributes ``+-0.25*jz`` to the diagonal and ``0.125*j`` off the diagonal, so an
undirected bond listed in both directions gives ``+-jz/2`` and ``j/4``."""
from ..operators import species_hops  # noqa: F401  (kept for API discoverability)
from .abc import AbstractSpinModel

__all__ = ["HeisenbergModel"]


class HeisenbergModel(AbstractSpinModel):
    def __init__(self, latt, j=1.0, jz=None):
        super().__init__(latt.num_sites)
        self.latt = latt
        self.j = j
        self.jz = j if jz is None else jz

    def _pairs(self):
        return [(p1, int(p2)) for p1 in range(self.num_sites) for p2 in self.latt.neighbors(p1)]

    def _operator_spec(self):
        if self.num_sites > 32:
            return None
        return dict(pairs=self._pairs(), j=self.j, jz=self.jz)

    def _hamiltonian_data(self, states):
        """COO triplets in the reference's emission order (state-major, then directed
        pairs): diagonal entry first, then the spin-flip entry when the two spins differ."""
        index = {int(s): k for k, s in enumerate(states)}
        pairs = self._pairs()
        quarter = 0.25
        for idx1, s1 in enumerate(states):
            s1 = int(s1)
            for pos1, pos2 in pairs:
                b1, b2 = (s1 >> pos1) & 1, (s1 >> pos2) & 1
                sign = (-1) ** b1 * (-1) ** b2
                yield idx1, idx1, sign * quarter * self.jz
                if b1 != b2:
                    yield idx1, index[s1 ^ (1 << pos1) ^ (1 << pos2)], quarter * self.j / 2
