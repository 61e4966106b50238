# -*- coding: utf-8 -*-
"""Model base classes behind cmpy's model interface (reference: cmpy/models/abc.py).

What a model author sees is unchanged: a model keeps its parameters as attributes *and* as
mapping items, implements the generator hook ``_hamiltonian_data(up_states, dn_states)``
(``(states)`` for spin models) yielding ``(row, col, value)`` and gets ``hamiltonian_data``,
``hamilton_operator`` and ``hamiltonian`` for free (reference: cmpy/models/abc.py:160-260).

What is different here is where ``hamilton_operator`` goes.  A model may also describe itself
through ``_operator_spec()`` (bond list, hop amplitudes, on-site and interaction energies); the
base class then builds the matrix-free GPU operator (kernels K4 / K5) and never materialises a
triplet.  The description is only trusted when it belongs to the generator hook in force: a
subclass that overrides ``_hamiltonian_data`` without also overriding ``_operator_spec`` gets
the triplet (COO) operator, as in the reference, so a changed Hamiltonian is never silently
replaced by its parent's.
"""
import json
from abc import ABC, abstractmethod
from collections.abc import MutableMapping

import numpy as np

from ..basis import Basis, SpinBasis
from ..operators import HamiltonOperator, SectorHamiltonOperator, SpinHamiltonOperator

__all__ = ["ModelParameters", "AbstractModel", "AbstractSpinModel", "AbstractManyBodyModel"]

_STORE = "_model_values"


class ModelParameters(MutableMapping):
    """Named parameters, readable and writable as ``model.u`` and as ``model["u"]``.

    The values live in one ordered dict kept under a private instance attribute; attribute
    access falls through to it only for names that were registered as parameters, so ordinary
    attributes (``basis``, ``neighbors`` ...) behave normally."""

    def __init__(self, **params):
        object.__setattr__(self, _STORE, dict(params))

    # ---- the mapping protocol -----------------------------------------------------------
    def _values(self):
        return self.__dict__[_STORE]

    def __getitem__(self, name):
        return self._values()[name]

    def __setitem__(self, name, value):
        self._values()[name] = value

    def __delitem__(self, name):
        del self._values()[name]

    def __iter__(self):
        return iter(self._values())

    def __len__(self):
        return len(self._values())

    # ---- attribute access to registered names ----------------------------------------------
    def __getattr__(self, name):
        # only reached when normal lookup failed
        store = self.__dict__.get(_STORE)
        if store is not None and name in store:
            return store[name]
        raise AttributeError(f"{type(self).__name__!s} has no attribute or parameter {name!r}")

    def __setattr__(self, name, value):
        store = self.__dict__.get(_STORE)
        if store is not None and name in store:
            store[name] = value
        else:
            object.__setattr__(self, name, value)

    # ---- the reference's convenience methods -------------------------------------------------
    @property
    def params(self):
        return self._values()

    def set_param(self, key, value):
        self[key] = value

    def delete_param(self, key):
        del self[key]

    def rename_param(self, key, new_key):
        self[new_key] = self._values().pop(key)

    def key(self, decimals=None, delim="; "):
        def show(v):
            return f"{v:.{decimals}f}" if decimals is not None and isinstance(v, (int, float)) else v

        return delim.join(f"{k}={show(v)}" for k, v in self.items())

    def json(self):
        return json.dumps(self._values())

    def pformat(self):
        return ", ".join(f"{k}={v}" for k, v in self.items())

    def __repr__(self):
        return f"{type(self).__name__}({self._values()!s})"

    def __str__(self):
        return self.pformat()


class AbstractModel(ModelParameters, ABC):
    def __init__(self, **params):
        ModelParameters.__init__(self, **params)

    def __str__(self):
        return f"{type(self).__name__}({self.pformat()})"

    def hamiltonian(self, *args, **kwargs):
        pass

    # ---- shared by the spin and the fermion flavour ------------------------------------------
    def _drain(self, *state_lists):
        """Runs the generator hook and splits its triplets into three lists."""
        rows, cols, vals = [], [], []
        for row, col, val in self._hamiltonian_data(*state_lists):
            rows.append(row)
            cols.append(col)
            vals.append(val)
        return rows, cols, vals

    def _trusted_spec(self):
        """``_operator_spec()`` if it describes the generator hook in force (see module docstring)."""
        mro = type(self).__mro__
        owner = {}
        for name in ("_operator_spec", "_hamiltonian_data"):
            owner[name] = next((i for i, c in enumerate(mro) if name in c.__dict__), len(mro))
        if owner["_operator_spec"] > owner["_hamiltonian_data"]:
            return None   # the hook was overridden further down the hierarchy than the description
        return self._operator_spec()


class AbstractSpinModel(AbstractModel):
    """Spin models on ``SpinBasis`` (reference: cmpy/models/abc.py:160-203)."""

    def __init__(self, num_sites=0, **params):
        super().__init__(**params)
        self.basis = SpinBasis()
        self.init_basis(num_sites)

    num_sites = property(lambda self: self.basis.num_sites)
    spins = property(lambda self: self.basis.spins)

    def init_basis(self, num_sites, init_sectors=None):
        self.basis.init(num_sites, init_sectors)

    def get_states(self, s=None):
        return self.basis.get_states(s)

    @abstractmethod
    def _hamiltonian_data(self, states):
        pass

    def _operator_spec(self):
        """Optional ``dict(pairs=[(pos1, pos2), ...], j=..., jz=...)``: enables the matrix-free
        GPU operator (kernel K5)."""
        return None

    def hamiltonian_data(self, states):
        rows, cols, vals = self._drain(states)
        return vals, (rows, cols)

    def hamilton_operator(self, s=None, states=None, dtype=None):
        spec = self._trusted_spec() if states is None else None
        if spec is not None:
            n_up = None if s is None else self.basis.num_up(s)
            return SpinHamiltonOperator(self.num_sites, n_up, spec["pairs"], spec["j"], spec["jz"], dtype=dtype)
        states = self.get_states(s) if states is None else states
        vals, indices = self.hamiltonian_data(states)
        return HamiltonOperator(len(states), vals, indices, dtype=dtype)

    def hamiltonian(self, s=None, states=None, dtype=None):
        return self.hamilton_operator(s, states, dtype).toarray()


class AbstractManyBodyModel(AbstractModel):
    """Fermionic lattice models on ``Basis`` (reference: cmpy/models/abc.py:206-260)."""

    def __init__(self, num_sites=0, **params):
        super().__init__(**params)
        self.basis = Basis()
        self.init_basis(num_sites)

    num_sites = property(lambda self: self.basis.num_sites)
    fillings = property(lambda self: self.basis.fillings)

    def init_basis(self, num_sites, init_sectors=None):
        self.basis.init(num_sites, init_sectors)

    def iter_fillings(self):
        return self.basis.iter_fillings()

    def iter_sectors(self):
        return self.basis.iter_sectors()

    def get_sector(self, n_up=None, n_dn=None):
        return self.basis.get_sector(n_up, n_dn)

    @abstractmethod
    def _hamiltonian_data(self, up_states, dn_states):
        pass

    def _operator_spec(self):
        """Optional ``dict(bonds, hops, eps, u, sign_width)``: enables the matrix-free GPU
        operator (kernel K4)."""
        return None

    def hamiltonian_data(self, up_states, dn_states):
        rows, cols, vals = self._drain(up_states, dn_states)
        return vals, np.array([rows, cols], dtype=np.int64)

    def hamilton_operator(self, n_up=None, n_dn=None, sector=None, dtype=None):
        sector = self.basis.get_sector(n_up, n_dn) if sector is None else sector
        up_states, dn_states = sector.up_states, sector.dn_states
        spec = self._trusted_spec()
        if spec is not None:
            return SectorHamiltonOperator(self.num_sites, up_states, dn_states, spec["bonds"], spec["hops"],
                                          spec["eps"], spec["u"], spec["sign_width"], dtype=dtype)
        vals, indices = self.hamiltonian_data(up_states, dn_states)
        return HamiltonOperator(len(up_states) * len(dn_states), vals, indices, dtype=dtype)

    def hamiltonian(self, n_up=None, n_dn=None, sector=None, dtype=None):
        return self.hamilton_operator(n_up, n_dn, sector, dtype).toarray()
