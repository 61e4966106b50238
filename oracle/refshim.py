# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference
(`/root/reference`, dylanljones/cmpy) import under numpy>=2 / scipy>=1.15 when
matplotlib, colorcet, lattpy, gftool are absent.

Used by `oracle/make_golden.py` (fixture generation, in the dev container only)
and by `tests/test_oracle_vs_reference.py` (skipped when the reference tree is
not present, e.g. on the GPU box).  Nothing in `cmpy_b200/` imports this.
"""
import os
import sys
import types
import collections
import collections.abc

REFERENCE_ROOT = os.environ.get("CMPY_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cmpy"))


class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any


class _Any(metaclass=_Meta):
    """Usable as base class, attribute bag and callable."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any

    def __call__(self, *a, **k):
        return _Any()


class _Mod(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any


_loaded = None


def load_reference():
    """Returns the reference `cmpy` package (imported from REFERENCE_ROOT)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np

    if not hasattr(np, "infty"):
        np.infty = np.inf  # matrix.py:44, exactdiag.py:146
    collections.Sequence = collections.abc.Sequence
    import scipy.sparse.linalg as _sla  # noqa
    import scipy.sparse.linalg.interface as _pub
    import scipy.sparse.linalg._interface as _priv

    _pub.IdentityOperator = _priv.IdentityOperator
    _pub.LinearOperator = _priv.LinearOperator
    for n in ["matplotlib", "matplotlib.colors", "matplotlib.pyplot", "colorcet",
              "gftool", "gftool.fourier"]:
        if n not in sys.modules:
            m = _Mod(n)
            m.__path__ = []
            sys.modules[n] = m
    if "lattpy" not in sys.modules:
        m = _Mod("lattpy")
        m.__path__ = []
        m.Lattice = type("Lattice", (), {})
        sys.modules["lattpy"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import cmpy  # noqa
    import cmpy.basis, cmpy.operators, cmpy.exactdiag, cmpy.greens, cmpy.models  # noqa

    _loaded = cmpy
    return cmpy


class ChainStandIn:
    """lattpy stand-in: open (or periodic) chain with `.num_sites`, `.neighbors(i)`,
    `.neighbor_pairs(unique)` (reference: models/heisenberg.py:14,20; hubbard.py:62-64)."""

    def __init__(self, num_sites, periodic=False):
        self.num_sites = num_sites
        self.periodic = periodic

    def neighbors(self, i):
        n = self.num_sites
        out = []
        if self.periodic and n > 2:
            out = [(i - 1) % n, (i + 1) % n]
        else:
            if i - 1 >= 0:
                out.append(i - 1)
            if i + 1 < n:
                out.append(i + 1)
        return out

    def neighbor_pairs(self, unique=True):
        pairs = []
        for i in range(self.num_sites):
            for j in self.neighbors(i):
                if not unique or i < j:
                    pairs.append([i, j])
        return pairs, [1.0] * len(pairs)
