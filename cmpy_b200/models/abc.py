# -*- coding: utf-8 -*-
"""Model base classes with cmpy's interface (reference: cmpy/models/abc.py).

The plugin hook is unchanged: a model implements ``_hamiltonian_data(up_states, dn_states)``
(or ``(states)`` for spin models) yielding ``(row, col, value)``; ``hamiltonian_data`` drains
it into COO arrays.  What changes is ``hamilton_operator``: models that also implement
``_operator_spec`` get a matrix-free GPU operator (no triplets), everything else falls
through to the COO ``HamiltonOperator`` (GPU COO mat-vec) exactly as in the reference.
"""
import json
from abc import ABC, abstractmethod
from collections import OrderedDict
from collections.abc import MutableMapping
from typing import Any, Dict, Iterator, List, Optional

import numpy as np

from ..basis import Basis, SpinBasis
from ..operators import HamiltonOperator, SectorHamiltonOperator, SpinHamiltonOperator

__all__ = ["ModelParameters", "AbstractModel", "AbstractSpinModel", "AbstractManyBodyModel"]


class ModelParameters(MutableMapping):
    """Parameters reachable both as attributes and as dict items
    (reference: cmpy/models/abc.py:21-133)."""

    def __init__(self, **params):
        MutableMapping.__init__(self)
        self.__params__ = OrderedDict(params)

    @property
    def params(self) -> Dict[str, Any]:
        return self.__params__

    def set_param(self, key: str, value: Any) -> None:
        self.__params__[key] = value

    def delete_param(self, key: str) -> None:
        del self.__params__[key]

    def rename_param(self, key: str, new_key: str) -> None:
        self.__params__[new_key] = self.__params__.pop(key)

    def __len__(self) -> int:
        return len(self.__params__)

    def __getitem__(self, key: str) -> Any:
        return self.__params__[key]

    def __setitem__(self, key: str, value: Any) -> None:
        self.__params__[key] = value

    def __delitem__(self, key: str) -> None:
        del self.__params__[key]

    def __iter__(self) -> Iterator[str]:
        return iter(self.__params__)

    def __getattr__(self, key: str) -> Any:
        key = str(key)
        if not key.startswith("__") and key in self.__dict__.get("__params__", {}):
            return self.__dict__["__params__"][key]
        return super().__getattribute__(key)

    def __setattr__(self, key: str, value: Any) -> None:
        key = str(key)
        params = self.__dict__.get("__params__")
        if params is not None and not key.startswith("__") and key in params:
            params[key] = value
        else:
            super().__setattr__(key, value)

    def key(self, decimals=None, delim="; "):
        parts = []
        for k, v in self.__params__.items():
            if decimals is not None and isinstance(v, (int, float)):
                v = f"{v:.{decimals}f}"
            parts.append(f"{k}={v}")
        return delim.join(parts)

    def json(self):
        return json.dumps(self.__params__)

    def pformat(self):
        return ", ".join(f"{k}={v}" for k, v in self.__params__.items())

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({str(self.__params__)})"

    def __str__(self) -> str:
        return self.pformat()


class AbstractModel(ModelParameters, ABC):
    def __init__(self, **params):
        ModelParameters.__init__(self, **params)
        ABC.__init__(self)

    def __str__(self) -> str:
        return f"{self.__class__.__name__}({ModelParameters.__str__(self)})"

    def hamiltonian(self, *args, **kwargs):
        pass


class AbstractSpinModel(AbstractModel):
    """Base class of spin models (reference: cmpy/models/abc.py:160-203)."""

    def __init__(self, num_sites: Optional[int] = 0, **params):
        super().__init__(**params)
        self.basis: SpinBasis = SpinBasis()
        self.init_basis(num_sites)

    @property
    def num_sites(self) -> int:
        return self.basis.num_sites

    @property
    def spins(self) -> List[int]:
        return self.basis.spins

    def init_basis(self, num_sites: int, init_sectors: bool = None):
        self.basis.init(num_sites, init_sectors)

    def get_states(self, s: float = None):
        return self.basis.get_states(s)

    @abstractmethod
    def _hamiltonian_data(self, states):
        pass

    def _operator_spec(self):
        """Optional: ``dict(pairs=[(pos1, pos2), ...], j=..., jz=...)`` enabling the
        matrix-free GPU operator."""
        return None

    def hamiltonian_data(self, states):
        rows, cols, data = list(), list(), list()
        for row, col, val in self._hamiltonian_data(states):
            rows.append(row)
            cols.append(col)
            data.append(val)
        return data, (rows, cols)

    def hamilton_operator(self, s=None, states=None, dtype=None):
        spec = self._operator_spec() if states is None else None
        if spec is not None:
            n_up = None if s is None else self.basis.num_up(s)
            return SpinHamiltonOperator(self.num_sites, n_up, spec["pairs"], spec["j"], spec["jz"],
                                        dtype=dtype)
        if states is None:
            states = self.get_states(s)
        data, indices = self.hamiltonian_data(states)
        return HamiltonOperator(len(states), data, indices, dtype=dtype)

    def hamiltonian(self, s=None, states=None, dtype=None):
        return self.hamilton_operator(s, states, dtype).toarray()


class AbstractManyBodyModel(AbstractModel):
    """Base class of fermionic lattice models (reference: cmpy/models/abc.py:206-260)."""

    def __init__(self, num_sites: Optional[int] = 0, **params):
        super().__init__(**params)
        self.basis: Basis = Basis()
        self.init_basis(num_sites)

    @property
    def num_sites(self) -> int:
        return self.basis.num_sites

    @property
    def fillings(self) -> List[int]:
        return self.basis.fillings

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
        """Optional: ``dict(bonds, hops, eps, u, sign_width)`` enabling the matrix-free GPU
        operator (kernel K4)."""
        return None

    def hamiltonian_data(self, up_states, dn_states):
        rows, cols, data = list(), list(), list()
        for row, col, val in self._hamiltonian_data(up_states, dn_states):
            rows.append(row)
            cols.append(col)
            data.append(val)
        return data, np.array([rows, cols], dtype=np.int64)

    def hamilton_operator(self, n_up=None, n_dn=None, sector=None, dtype=None):
        if sector is None:
            sector = self.basis.get_sector(n_up, n_dn)
        up_states, dn_states = sector.up_states, sector.dn_states
        spec = self._operator_spec()
        if spec is not None:
            return SectorHamiltonOperator(self.num_sites, up_states, dn_states, spec["bonds"],
                                          spec["hops"], spec["eps"], spec["u"], spec["sign_width"],
                                          dtype=dtype)
        size = len(up_states) * len(dn_states)
        data, indices = self.hamiltonian_data(up_states, dn_states)
        return HamiltonOperator(size, data, indices, dtype=dtype)

    def hamiltonian(self, n_up=None, n_dn=None, sector=None, dtype=None):
        return self.hamilton_operator(n_up, n_dn, sector, dtype).toarray()
