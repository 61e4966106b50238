# -*- coding: utf-8 -*-
"""Single impurity Anderson model in star geometry (reference: cmpy/models/anderson.py).

Site 0 is the impurity, sites 1..N_b the bath.  The reference projects the hybridisation
with fermion-sign width 0, i.e. signless hops (anderson.py:149,158); ``mu = u/2`` when
``None`` (anderson.py:61)."""
import numpy as np

from ..basis import UP
from ..operators import project_hubbard_inter, project_onsite_energy, project_hopping
from .abc import AbstractManyBodyModel

__all__ = ["SingleImpurityAndersonModel"]


class SingleImpurityAndersonModel(AbstractManyBodyModel):
    def __init__(self, u=2.0, eps_imp=0.0, eps_bath=0.0, v=1.0, mu=None, temp=0.0):
        mu = u / 2 if mu is None else mu
        eps_bath = np.atleast_1d(eps_bath)
        v = np.atleast_1d(v)
        if len(eps_bath) > 1 and len(v) == 1:
            v = np.ones(len(eps_bath)) * v[0]
        if len(eps_bath) == 1 and len(v) > 1:
            eps_bath = np.ones(len(v)) * eps_bath[0]
        assert len(eps_bath) == len(v), (
            f"Shape of bath on-site energy {len(eps_bath)} doesn't match hybridization {len(v)}!")
        super().__init__(len(eps_bath) + 1, u=u, eps_imp=eps_imp, eps_bath=eps_bath, v=v, mu=mu,
                         temp=temp)

    @property
    def num_bath(self) -> int:
        return len(self.eps_bath)

    @property
    def num_sites(self) -> int:
        return self.num_bath + 1

    @property
    def beta(self) -> float:
        return 1 / self.temp

    def update_bath_energy(self, eps_bath) -> None:
        eps_bath = np.atleast_1d(eps_bath).astype(np.float64)
        if eps_bath.shape[0] != self.num_bath:
            raise ValueError(f"Dimension of the new bath energy {eps_bath.shape} "
                             f"does not match number of baths {self.num_bath}")
        self.eps_bath = eps_bath  # noqa

    def update_hybridization(self, v) -> None:
        v = np.atleast_1d(v).astype(np.float64)
        if v.shape[0] != self.num_bath:
            raise ValueError(f"Dimension of the new hybridization {v.shape} "
                             f"does not match number of baths {self.num_bath}")
        self.v = v  # noqa

    def hybridization_func(self, z: np.ndarray) -> np.ndarray:
        """Delta(z) = sum_i |V_i|^2 / (z - eps_i) (reference: anderson.py:127-145)."""
        x = np.asarray(z)[..., np.newaxis]
        return np.sum(np.square(np.abs(self.v)) / (x - self.eps_bath), axis=-1)

    def _site_arrays(self):
        u = np.append(self.u, np.zeros(self.num_bath))
        eps = np.append(self.eps_imp - self.mu, self.eps_bath)
        return np.asarray(u, dtype=np.float64), np.asarray(eps, dtype=np.float64)

    def _hamiltonian_data(self, up_states, dn_states):
        u, eps = self._site_arrays()
        yield from project_onsite_energy(up_states, dn_states, eps)
        yield from project_hubbard_inter(up_states, dn_states, u)
        for j in range(self.num_bath):
            yield from project_hopping(up_states, dn_states, 0, 0, j + 1, self.v[j])

    def _operator_spec(self):
        u, eps = self._site_arrays()
        return dict(bonds=[(0, j + 1) for j in range(self.num_bath)],
                    hops=np.asarray(self.v, dtype=np.float64), eps=eps, u=u, sign_width=0)

    def impurity_gf0(self, z):
        return 1 / (z + self.mu + self.eps_imp - self.hybridization_func(z))

    def impurity_gf(self, z, sigma=UP):
        from ..exactdiag import gf_lehmann

        return gf_lehmann(self, z, beta=1 / self.temp, pos=0, sigma=sigma).gf

    def pformat(self):
        return (f"U={self.u}, ε_i={self.eps_imp}, ε_b={self.eps_bath}, v={self.v}, "
                f"μ={self.mu}, T={self.temp}")
