# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/hv_oracle.c (the C port of the
reference's H.v used as checker at sizes numpy cannot reach and as the CPU baseline of
bench.py).  Never imported by cmpy_b200/."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int8, c_int32, c_int64

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hv_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libhv_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(OUT_DIR, exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC],
                       check=True)
    return OUT


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(build())
        except OSError:  # built for another CPU: rebuild here
            _lib = ctypes.CDLL(build(force=True))
        _lib.orc_enumerate.restype = c_int64
        _lib.orc_enumerate.argtypes = [c_int, c_int, POINTER(c_int64), c_int64]
        _lib.orc_species_hops.restype = None
        _lib.orc_species_hops.argtypes = [POINTER(c_int64), c_int64, c_int, c_int, c_int,
                                          POINTER(c_int32), POINTER(c_int8)]
        _lib.orc_hubbard_hv_rows.restype = None
        _lib.orc_hubbard_hv_rows.argtypes = [
            c_int, POINTER(c_int64), c_int64, POINTER(c_int64), c_int64, c_int, POINTER(c_double),
            POINTER(c_double), POINTER(c_double), POINTER(c_int32), POINTER(c_int8), POINTER(c_int32),
            POINTER(c_int8), POINTER(c_double), POINTER(c_double), c_int64, c_int64, c_int]
        _lib.orc_max_threads.restype = c_int
        _lib.orc_heisenberg_hv_range.restype = None
        _lib.orc_heisenberg_hv_range.argtypes = [c_int, c_int, POINTER(c_int32), POINTER(c_int32), c_double, c_double,
                                                 POINTER(c_double), POINTER(c_double), c_int64, c_int64, c_int]
        _lib.orc_binomial.restype = c_int64
        _lib.orc_binomial.argtypes = [c_int, c_int]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(POINTER(t))


def enumerate_states(num_sites, n):
    cnt = lib().orc_enumerate(num_sites, n, None, 0)
    out = np.empty(cnt, dtype=np.int64)
    lib().orc_enumerate(num_sites, n, _p(out, c_int64), cnt)
    return out


def max_threads():
    return int(lib().orc_max_threads())


class HubbardOracle:
    """Precomputed hop tables (the analogue of the reference's COO build) + matrix-free H.v."""

    def __init__(self, num_sites, up_states, dn_states, bonds, hops, eps, u, width):
        self.num_sites = int(num_sites)
        self.up = np.ascontiguousarray(up_states, dtype=np.int64)
        self.dn = np.ascontiguousarray(dn_states, dtype=np.int64)
        self.bonds = [(int(i), int(j)) for i, j in bonds]
        self.hops = np.ascontiguousarray(hops, dtype=np.float64)
        self.eps = np.ascontiguousarray(eps, dtype=np.float64)
        self.u = np.ascontiguousarray(u, dtype=np.float64)
        nb = len(self.bonds)
        self.tgt_up = np.empty((max(nb, 1), len(self.up)), dtype=np.int32)
        self.sgn_up = np.empty((max(nb, 1), len(self.up)), dtype=np.int8)
        self.tgt_dn = np.empty((max(nb, 1), len(self.dn)), dtype=np.int32)
        self.sgn_dn = np.empty((max(nb, 1), len(self.dn)), dtype=np.int8)
        for b, (i, j) in enumerate(self.bonds):
            lib().orc_species_hops(_p(self.up, c_int64), len(self.up), int(width), i, j,
                                   _p(self.tgt_up[b], c_int32), _p(self.sgn_up[b], c_int8))
            lib().orc_species_hops(_p(self.dn, c_int64), len(self.dn), int(width), i, j,
                                   _p(self.tgt_dn[b], c_int32), _p(self.sgn_dn[b], c_int8))

    @property
    def size(self):
        return len(self.up) * len(self.dn)

    def matvec_rows(self, x, row0=0, nrows=None, nthreads=0):
        nrows = len(self.up) - row0 if nrows is None else nrows
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.size
        y = np.empty(nrows * len(self.dn), dtype=np.float64)
        lib().orc_hubbard_hv_rows(
            self.num_sites, _p(self.up, c_int64), len(self.up), _p(self.dn, c_int64), len(self.dn),
            len(self.bonds), _p(self.hops, c_double), _p(self.eps, c_double), _p(self.u, c_double),
            _p(self.tgt_up, c_int32), _p(self.sgn_up, c_int8), _p(self.tgt_dn, c_int32),
            _p(self.sgn_dn, c_int8), _p(x, c_double), _p(y, c_double), int(row0), int(nrows),
            int(nthreads))
        return y

    def matvec(self, x, nthreads=0):
        return self.matvec_rows(x, 0, None, nthreads)


def hubbard_oracle(num_sites, n_up, n_dn, neighbors, inter, eps, hop, width=None):
    """Hubbard model with scalar parameters (eps already eps-mu), bonds with i<j only."""
    bonds = [(i, j) for i, j in neighbors if i < j]
    up = enumerate_states(num_sites, n_up)
    dn = enumerate_states(num_sites, n_dn)
    return HubbardOracle(num_sites, up, dn, bonds, np.full(len(bonds), hop), np.full(num_sites, eps),
                         np.full(num_sites, inter), num_sites if width is None else width)


class HeisenbergOracle:
    """Matrix-free H.v of `HeisenbergModel._hamiltonian_data` (cmpy/models/heisenberg.py:19-40) on the sector
    of `n_up` up spins, any contiguous range of rows; `neighbor_lists[pos1]` = the reference's
    `latt.neighbors(pos1)` (directed pairs)."""

    def __init__(self, num_sites, n_up, neighbor_lists, j=1.0, jz=1.0):
        self.num_sites, self.n_up, self.j, self.jz = int(num_sites), int(n_up), float(j), float(jz)
        ptr = [0]
        idx = []
        for pos1 in range(self.num_sites):
            idx.extend(int(p) for p in neighbor_lists[pos1])
            ptr.append(len(idx))
        self.ptr = np.asarray(ptr, dtype=np.int32)
        self.idx = np.asarray(idx if idx else [0], dtype=np.int32)
        self.size = int(lib().orc_binomial(self.num_sites, self.n_up))

    def matvec_range(self, x, i0=0, count=None, nthreads=0):
        count = self.size - i0 if count is None else count
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.size and 0 <= i0 and i0 + count <= self.size
        y = np.empty(count, dtype=np.float64)
        lib().orc_heisenberg_hv_range(self.num_sites, self.n_up, _p(self.ptr, c_int32), _p(self.idx, c_int32),
                                      self.j, self.jz, _p(x, c_double), _p(y, c_double), int(i0), int(count),
                                      int(nthreads))
        return y

    def matvec(self, x, nthreads=0):
        return self.matvec_range(x, 0, None, nthreads)
