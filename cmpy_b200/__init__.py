# -*- coding: utf-8 -*-
"""cmpy_b200 -- B200-native (sm_100a) exact-diagonalisation engine behind cmpy's Python
surface.  ``import cmpy_b200 as cmpy`` is the intended drop-in for the accelerated hot path:
``Basis/get_sector``, the ``project_*`` projectors, ``HamiltonOperator.matvec``, the Hubbard,
Heisenberg and Anderson models and the ``exactdiag`` / ``greens`` entry points.

All compute goes through ``libcmpy_b200.so`` (hand-written CUDA, C ABI in
``include/cmpy_b200.h``); there is no CPU fallback."""
import logging

logger = logging.getLogger("cmpy")  # same logger name as the reference (cmpy/_utils.py:15-29)

from .basis import (  # noqa: E402
    UP, DN, SPIN_CHARS, state_label, binstr, binarr, binidx, overlap, occupations, create,
    annihilate, SpinState, State, Sector, Basis, SpinBasis,
)
from .matrix import EigenState, is_hermitian  # noqa: E402
from .operators import (  # noqa: E402
    project_up, project_dn, project_elements_up, project_elements_dn, project_onsite_energy,
    project_hubbard_inter, project_hopping, LinearOperator, HamiltonOperator,
    SectorHamiltonOperator, SpinHamiltonOperator, TimeEvolutionOperator, CreationOperator,
    AnnihilationOperator,
)
from .models.abc import ModelParameters, AbstractModel, AbstractManyBodyModel  # noqa: E402
from . import models, exactdiag, greens, dmft, observables  # noqa: E402

__version__ = "0.1.0"
