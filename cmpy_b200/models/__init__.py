# -*- coding: utf-8 -*-
"""Model classes of the accelerated hot path (reference: cmpy/models/__init__.py).
Ising and tight-binding models are out of scope."""
from .abc import ModelParameters, AbstractModel, AbstractSpinModel, AbstractManyBodyModel
from .hubbard import HubbardModel, hubbard_hamiltonian
from .heisenberg import HeisenbergModel
from .anderson import SingleImpurityAndersonModel
