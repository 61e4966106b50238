# -*- coding: utf-8 -*-
"""The two items of cmpy/matrix.py that the hot path touches:
``EigenState`` (return type of ``compute_groundstate``, reference: cmpy/matrix.py:42-47)
and ``is_hermitian`` (test helper, reference: cmpy/matrix.py:191-212).  The dense
decomposition / plotting utilities of that module are out of scope."""
from typing import NamedTuple

import numpy as np

__all__ = ["EigenState", "is_hermitian"]


class EigenState(NamedTuple):
    energy: float = np.inf
    state: np.ndarray = None
    n_up: int = None
    n_dn: int = None


def is_hermitian(a, rtol: float = 1e-05, atol: float = 1e-08) -> bool:
    a = np.asarray(a)
    return bool(np.allclose(a, np.conj(a).T, rtol=rtol, atol=atol))
