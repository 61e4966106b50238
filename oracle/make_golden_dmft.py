# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/reference_dmft.npz` from the UNMODIFIED
reference's two-site DMFT loop (cmpy/dmft/twosite.py:85-178), the caller of the impurity G(z)
path.   NUMBA_DISABLE_JIT=1 python oracle/make_golden_dmft.py
"""
import os
import sys
import types

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "reference_dmft.npz")


def main():
    refshim.load_reference()
    import numba  # noqa: F401  (imported before the colorama stand-in so numba does not pick it up)

    col = types.ModuleType("colorama")

    class _C:
        def __getattr__(self, k):
            return ""

    col.Fore, col.Style, col.Back, col.init = _C(), _C(), _C(), (lambda *a, **k: None)
    sys.modules["colorama"] = col
    from cmpy.dmft import twosite
    from cmpy.dmft.utils import self_energy, quasiparticle_weight, bethe_gf_omega

    z = np.linspace(-6, 6, 2001) + 1e-2j
    g = {"z": z}
    us = np.array([0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 5.9, 6.5])
    g["u_ref"] = us
    g["v_ref"] = np.array([twosite.twosite_dmft_half_filling(z, u, t=1.0, verbose=False, ref=True).v[0] for u in us])
    us2 = np.array([2.0, 4.0])
    g["u_ed"] = us2
    vs = []
    for u in us2:
        siam = twosite.twosite_dmft_half_filling(z, u, t=1.0, beta=50.0, verbose=False, ref=False, max_iter=100)
        vs.append(siam.v[0])
        g[f"gf_latt_u{int(u)}"] = twosite.compute_lattice_greens_function(z, siam, 1.0, ref=False)
    g["v_ed"] = np.array(vs)
    g["params_ref_u4_v07"] = np.array(twosite.impurity_params_ref(4.0, 0.7))
    g["gf_ref_u4_v07"] = twosite.impurity_gf_ref(z, 4.0, 0.7)
    sig = self_energy(1 / (z + 0.3), twosite.impurity_gf_ref(z, 4.0, 0.7))
    g["sigma_test"] = sig
    g["qp_test"] = np.float64(quasiparticle_weight(z.real, sig, thresh=1e-10))
    g["bethe"] = bethe_gf_omega(z, 1.0)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in g.items()})
    print(g["v_ref"], g["v_ed"])


if __name__ == "__main__":
    main()
