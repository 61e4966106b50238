# -*- coding: utf-8 -*-
"""Fock-basis containers with cmpy's interface (reference: cmpy/basis.py).

Bit ``i`` of a spin string is site ``i`` (LSB = site 0).  A sector is the product of all
spin-up strings with ``n_up`` particles and all spin-down strings with ``n_dn`` particles,
both in ascending integer order; the composite index is ``up_idx * num_dn + dn_idx``
(reference: cmpy/operators.py:33-90, cmpy/basis.py:568-575).

Sector enumeration does not use the reference's ``itertools.permutations`` walk
(cmpy/basis.py:655-666, O(L!)): the ascending fixed-popcount list is produced by the
combinadic-unranking CUDA kernel ``cmpy_sector_enumerate`` (csrc/sector.cuh).
"""
from collections import defaultdict
from itertools import product
from typing import Iterable, List, Tuple, Union

import numpy as np

from . import _lib

__all__ = [
    "UP", "DN", "SPIN_CHARS", "state_label", "bit_count", "binstr", "binarr", "binidx",
    "get_ibit", "set_ibit", "overlap", "occupations", "create", "annihilate", "SpinState",
    "State", "Sector", "Basis", "SpinBasis", "spinstate_label", "upper_sector", "lower_sector",
    "enumerate_states", "rank_states",
]

UP, DN = 1, 2  # reference: cmpy/basis.py:40

EMPTY_CHAR, UP_CHAR, DN_CHAR, UD_CHAR = ".", "↑", "↓", "⇅"
SPIN_CHARS = {0: EMPTY_CHAR, UP: UP_CHAR, DN: DN_CHAR, 3: UD_CHAR}


# -- device-side enumeration / ranking (K1) ------------------------------------------------

def enumerate_states_device(num_sites: int, n: int):
    """Ascending ``num_sites``-bit integers with ``n`` set bits as a CUDA int64 tensor."""
    torch = _lib.require_cuda()
    if not 0 <= n <= num_sites:
        raise ValueError(f"filling {n} not realisable with {num_sites} sites")
    count = _lib.binomial(num_sites, n)
    out = torch.empty(count, dtype=torch.int64, device=_lib.device())
    _lib.check(_lib.lib().cmpy_sector_enumerate(num_sites, n, _lib.ptr(out), _lib.stream_ptr()),
               "cmpy_sector_enumerate")
    return out


def enumerate_states(num_sites: int, n: int) -> np.ndarray:
    """Host copy of :func:`enumerate_states_device` (np.int64)."""
    return enumerate_states_device(num_sites, n).cpu().numpy()


def rank_states(states) -> np.ndarray:
    """Combinadic (colex) rank of each state inside the ascending list of states with the
    same popcount -- what ``bisect_left(states, x)`` returns in the reference
    (cmpy/operators.py:276-299). Accepts a numpy array or CUDA tensor; returns the same kind."""
    torch = _lib.require_cuda()
    is_tensor = isinstance(states, torch.Tensor)
    st = states if is_tensor else torch.as_tensor(np.ascontiguousarray(states, dtype=np.int64))
    st = st.to(device=_lib.device(), dtype=torch.int64).contiguous()
    out = torch.empty_like(st)
    _lib.check(_lib.lib().cmpy_sector_rank(_lib.ptr(st), st.numel(), _lib.ptr(out),
                                           _lib.stream_ptr()), "cmpy_sector_rank")
    return out if is_tensor else out.cpu().numpy()


# -- bit helpers (pure host arithmetic) ------------------------------------------------------

def bit_count(num: int) -> int:
    """Number of set bits."""
    return int(num).bit_count()


def binstr(num: int, width: int = 0) -> str:
    """Binary string of ``num`` (site 0 is the right-most character), zero padded."""
    width = width or 0
    return format(int(num), "b").rjust(width, "0")


def binarr(num: int, width: int = None, dtype=None) -> np.ndarray:
    """Bits of ``num`` as an array, element ``i`` = site ``i``."""
    num = int(num)
    nbits = max(width or 0, num.bit_length(), 1)
    return np.array([(num >> i) & 1 for i in range(nbits)], dtype=dtype or np.int64)


def binidx(num, width: int = None) -> Iterable[int]:
    """Indices of the set bits, ascending."""
    num = int(num)
    return [i for i in range(num.bit_length()) if (num >> i) & 1]


def get_ibit(num: int, index: int, length: int = 1) -> int:
    return (num >> index) & ((1 << length) - 1)


def set_ibit(num: int, index: int, value: int, length: int = 1) -> int:
    field = ((1 << length) - 1) << index
    return (num & ~field) | ((value << index) & field)


def overlap(num1: int, num2: int, width: int = None, dtype=None) -> np.ndarray:
    return binarr(num1 & num2, width, dtype)


def occupations(num: int, width: int = None, dtype=None) -> np.ndarray:
    return binarr(num, width, dtype)


def create(num: int, pos: int) -> Union[int, None]:
    """State with a particle added at ``pos``; ``None`` when the site is occupied."""
    bit = 1 << pos
    return None if num & bit else num | bit


def annihilate(num: int, pos: int) -> Union[int, None]:
    """State with the particle at ``pos`` removed; ``None`` when the site is empty."""
    bit = 1 << pos
    return num & ~bit if num & bit else None


def state_label(up_num: int, dn_num: int, digits: int = None) -> str:
    """One character per site: ``.``, ``↑``, ``↓`` or ``⇅``; site 0 first."""
    nchar = max(int(up_num).bit_length(), int(dn_num).bit_length(), 1, digits or 0)
    return "".join(SPIN_CHARS[((up_num >> i) & 1) + 2 * ((dn_num >> i) & 1)] for i in range(nchar))


def spinstate_label(state: int, digits: int = None) -> str:
    return "".join(UP_CHAR if c == "1" else DN_CHAR for c in binstr(state, digits))


class SpinState(int):
    """Integer with occupation helpers (reference: cmpy/basis.py:422-470)."""

    @property
    def n(self) -> int:
        return int(self).bit_count()

    def binstr(self, width: int = None) -> str:
        return binstr(self, width)

    def binarr(self, width: int = None, dtype=None) -> np.ndarray:
        return binarr(self, width, dtype)

    def occ(self, pos: int) -> int:
        return self & (1 << pos)

    def occupations(self, dtype=None) -> np.ndarray:
        return binarr(self, dtype=dtype)

    def overlap(self, other, dtype=None) -> np.ndarray:
        return overlap(self, other, dtype=dtype)

    def create(self, pos: int):
        new = create(self, pos)
        return None if new is None else type(self)(new)

    def annihilate(self, pos: int):
        new = annihilate(self, pos)
        return None if new is None else type(self)(new)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.binstr()})"

    def __str__(self) -> str:
        return self.binstr()


class State:
    """A spin-up string paired with a spin-down string."""

    __slots__ = ["up", "dn", "num_sites"]

    def __init__(self, up, dn, num_sites: int = None):
        self.up = SpinState(up)
        self.dn = SpinState(dn)
        self.num_sites = num_sites

    def label(self, width: int = None) -> str:
        return state_label(self.up, self.dn, self.num_sites if width is None else width)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.up}, {self.dn})"

    def __str__(self) -> str:
        return f"{type(self).__name__}: {self.label()}"

    def __eq__(self, other) -> bool:
        return self.up == other.up and self.dn == other.dn


def upper_sector(n_up: int, n_dn: int, sigma: int, num_sites: int) -> Union[Tuple[int, int], None]:
    """Fillings with one more ``sigma`` particle, or ``None`` when the band is full."""
    if sigma == UP:
        return (n_up + 1, n_dn) if n_up < num_sites else None
    if sigma == DN:
        return (n_up, n_dn + 1) if n_dn < num_sites else None
    return None


def lower_sector(n_up: int, n_dn: int, sigma: int) -> Union[Tuple[int, int], None]:
    """Fillings with one ``sigma`` particle less, or ``None`` when there is none."""
    if sigma == UP:
        return (n_up - 1, n_dn) if n_up > 0 else None
    if sigma == DN:
        return (n_up, n_dn - 1) if n_dn > 0 else None
    return None


class Sector:
    """Spin-up and spin-down string lists of one (n_up, n_dn) block."""

    __slots__ = ["num_sites", "n_up", "n_dn", "up_states", "dn_states"]

    def __init__(self, up_states, dn_states, n_up=None, n_dn=None, num_sites=0):
        self.num_sites = num_sites
        self.n_up = n_up
        self.n_dn = n_dn
        self.up_states = up_states
        self.dn_states = dn_states

    @property
    def states(self):
        for up, dn in product(self.up_states, self.dn_states):  # up-major
            yield State(up, dn, num_sites=self.num_sites)

    @property
    def num_up(self) -> int:
        return len(self.up_states)

    @property
    def num_dn(self) -> int:
        return len(self.dn_states)

    @property
    def size(self) -> int:
        return self.num_up * self.num_dn

    @property
    def filling(self):
        return self.n_up, self.n_dn

    def state_labels(self):
        return [s.label(self.num_sites) for s in self.states]

    def __iter__(self):
        return iter(self.states)

    def __repr__(self):
        return (f"{type(self).__name__}(size: {self.size}, num_sites: {self.num_sites}, "
                f"filling: [{self.n_up}, {self.n_dn}])")

    def __str__(self):
        return f"{type(self).__name__}({self.n_up}, {self.n_dn}, size: {self.size})"


class Basis:
    """All fermionic basis states of ``num_sites`` sites, organised by filling.

    ``get_states(n)`` keeps the reference's return types (cmpy/basis.py:655-666): python
    lists for ``n in (None, 0, 1)``, ``np.ndarray[int64]`` otherwise, cached per filling.
    """

    __slots__ = ["size", "num_sites", "num_spinstates", "sectors", "fillings"]

    def __init__(self, num_sites: int = 0, init_sectors: bool = False):
        self.size = 0
        self.num_sites = 0
        self.num_spinstates = 0
        self.sectors = defaultdict(list)
        self.fillings = [0]
        self.init(num_sites, init_sectors)

    def init(self, num_sites: int, init_sectors: bool = False):
        self.num_sites = num_sites
        self.num_spinstates = 1 << num_sites
        self.size = self.num_spinstates ** 2
        self.sectors = defaultdict(list)
        self.fillings = list(range(num_sites + 1))
        if init_sectors:
            for state in range(self.num_spinstates):
                self.sectors[state.bit_count()].append(state)

    def generate_states(self, n: int = None) -> Union[list, np.ndarray]:
        if n is None:
            return list(range(self.num_spinstates))
        if n == 0:
            return [0]
        if n == 1:
            return [1 << site for site in range(self.num_sites)]
        return enumerate_states(self.num_sites, n)

    def get_states(self, n: int = None) -> List[int]:
        if n not in self.sectors:
            self.sectors[n] = self.generate_states(n)
        return self.sectors[n]

    def get_sector(self, n_up: int = None, n_dn: int = None) -> Sector:
        return Sector(self.get_states(n_up), self.get_states(n_dn), n_up, n_dn, self.num_sites)

    def iter_fillings(self):
        return product(self.fillings, repeat=2)

    def iter_sectors(self):
        for n_up, n_dn in self.iter_fillings():
            yield self.get_sector(n_up, n_dn)

    def keys(self):
        return self.fillings

    def check(self, n_up, n_dn) -> bool:
        return n_up in self.fillings and n_dn in self.fillings

    def upper_sector(self, n_up, n_dn, sigma):
        fill = upper_sector(n_up, n_dn, sigma, self.num_sites)
        return None if fill is None else self.get_sector(*fill)

    def lower_sector(self, n_up, n_dn, sigma):
        fill = lower_sector(n_up, n_dn, sigma)
        return None if fill is None else self.get_sector(*fill)

    def __getitem__(self, item):
        if hasattr(item, "__len__"):
            return self.get_sector(*item)
        return self.get_states(item)

    def __repr__(self):
        return (f"{type(self).__name__}(size: {self.size}, num_sites: {self.num_sites}, "
                f"fillings: {self.fillings})")


class SpinBasis:
    """Spin-1/2 basis states by total magnetisation (reference: cmpy/basis.py:726-787).

    ``get_states(s)`` returns the ascending list of ``num_sites``-bit integers with
    ``num_sites/2 + s`` up spins; ``ValueError`` when ``s`` cannot be realised.  The
    reference filters ``range(2**N)`` in Python; here the list comes from the CUDA
    enumeration kernel.  (For N >= 28 do not materialise the list: build the operator with
    ``HeisenbergModel.hamilton_operator(s=...)``, which never needs it.)
    """

    __slots__ = ["size", "num_sites", "num_spinstates", "sectors", "spins"]

    def __init__(self, num_sites: int = 0):
        self.size = 0
        self.num_sites = 0
        self.num_spinstates = 0
        self.sectors = defaultdict(list)
        self.spins = list()
        self.init(num_sites)

    def init(self, num_sites: int, init_sectors: bool = False):
        self.num_sites = num_sites
        self.num_spinstates = 1 << num_sites
        self.size = num_sites ** 2  # (sic) reference: cmpy/basis.py:742
        self.sectors = defaultdict(list)
        self.spins = list(np.arange(-0.5 * num_sites, +0.5 * num_sites + 0.1, 1))
        if init_sectors:
            for state in range(self.num_spinstates):
                ones = state.bit_count()
                self.sectors[0.5 * (ones - (max(state.bit_length(), 1) - ones))].append(state)

    def num_up(self, s: float) -> int:
        """Number of up spins for total spin ``s``; ``ValueError`` if unrealisable."""
        n_up = self.num_sites / 2 + s
        n_dn = self.num_sites / 2 - s
        if (n_up % 1 != 0.0) or (n_dn % 1 != 0.0):
            raise ValueError(f"Total spin of {s} not realizable with {self.num_sites} sites")
        return int(n_up)

    def generate_states(self, s: float = None) -> List[int]:
        if s is None:
            return list(range(self.num_spinstates))
        n_up = self.num_up(s)
        if n_up < 0 or n_up > self.num_sites:
            return []
        return [int(v) for v in enumerate_states(self.num_sites, n_up)]

    def get_states(self, s: float = None) -> List[int]:
        if s not in self.sectors:
            self.sectors[s] = list(self.generate_states(s))
        return self.sectors[s]

    def state_labels(self, s: float = None):
        return [spinstate_label(state, self.num_sites) for state in self.get_states(s)]

    def __getitem__(self, item) -> List[int]:
        return self.get_states(item)

    def __repr__(self):
        return (f"{type(self).__name__}(size: {self.size}, num_sites: {self.num_sites}, "
                f"spins: {self.spins})")
