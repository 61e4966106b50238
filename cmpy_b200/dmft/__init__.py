# -*- coding: utf-8 -*-
"""Two-site DMFT at half filling: the reference's end-to-end caller of the impurity G(z) path
(reference: cmpy/dmft/twosite.py, cmpy/dmft/utils.py)."""
from .twosite import (impurity_params_ref, impurity_gf_ref, impurity_gf0, compute_impurity_gf,  # noqa: F401
                      compute_self_energy, twosite_dmft_half_filling, compute_lattice_greens_function)
from .utils import IterationStats, mix_values, self_energy, bethe_gf_omega, quasiparticle_weight  # noqa: F401
