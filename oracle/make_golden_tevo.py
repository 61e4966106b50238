# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/reference_tevo.npz`: outputs of the
UNMODIFIED reference's real-time Green's functions (cmpy/exactdiag.py:248-308, which drive
cmpy/linalg/expm_multiply.py through the reference HamiltonOperator) for small Hubbard chains.

    NUMBA_DISABLE_JIT=1 python oracle/make_golden_tevo.py        (dev container only)
"""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "reference_tevo.npz")


def main():
    refshim.load_reference()
    from cmpy.models import HubbardModel
    from cmpy import exactdiag as ed
    from cmpy.basis import UP

    g = {}
    for L, pos, stop, num in [(4, 0, 8.0, 81), (4, 2, 5.0, 51), (5, 0, 6.0, 61), (6, 0, 10.0, 101)]:
        nb = [[i, i + 1] for i in range(L - 1)]
        model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
        gs = ed.compute_groundstate(model)
        key = f"L{L}_p{pos}"
        g[key + "_gs_energy"] = np.float64(gs.energy)
        g[key + "_gs_sector"] = np.array([gs.n_up, gs.n_dn])
        g[key + "_gs_state"] = np.asarray(gs.state, dtype=np.float64)
        t, gg = ed.gf_greater(model, gs, 0.0, stop, num, pos, UP)
        t2, gl = ed.gf_lesser(model, gs, 0.0, stop, num, pos, UP)
        g[key + "_times"] = np.asarray(t)
        g[key + "_greater"] = np.asarray(gg, dtype=np.complex128)
        g[key + "_lesser"] = np.asarray(gl, dtype=np.complex128)
        print(key, gs.energy, gs.n_up, gs.n_dn, abs(gg[0]), abs(gl[0]))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays")


if __name__ == "__main__":
    main()
