"""Two-site DMFT caller against the unmodified reference's outputs
(tests/golden/reference_dmft.npz, oracle/make_golden_dmft.py)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_dmft.npz")


@pytest.mark.filterwarnings("ignore")
def test_dmft_closed_form_path_cpu():
    """ref=True never touches the device: poles/residues, self energy, quasiparticle weight, Bethe
    G and the converged hybridisation for 8 values of U, bit for bit."""
    from cmpy_b200 import dmft

    g = np.load(GOLD)
    z = g["z"]
    assert np.array_equal(np.array(dmft.impurity_params_ref(4.0, 0.7)), g["params_ref_u4_v07"])
    assert np.array_equal(dmft.impurity_gf_ref(z, 4.0, 0.7), g["gf_ref_u4_v07"])
    sig = dmft.self_energy(1 / (z + 0.3), dmft.impurity_gf_ref(z, 4.0, 0.7))
    assert np.array_equal(sig, g["sigma_test"])
    assert dmft.quasiparticle_weight(z.real, sig, thresh=1e-10) == float(g["qp_test"])
    assert np.array_equal(dmft.bethe_gf_omega(z, 1.0), g["bethe"])
    v = [dmft.twosite_dmft_half_filling(z, u, t=1.0, verbose=False, ref=True).v[0] for u in g["u_ref"]]
    assert np.array_equal(np.array(v), g["v_ref"])
    assert dmft.mix_values(1.0, 2.0, 0.25) == 1.25 and dmft.mix_values(1.0, 2.0) == 2.0
    with pytest.raises(AssertionError):
        dmft.mix_values(1.0, 2.0, 1.5)


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore")
def test_dmft_ed_solver_gpu():
    """ref=False: the impurity G(z) comes from the GPU Lehmann path (beta=50) -- converged V and
    the lattice G(z) against the reference; beta=inf uses the GPU continued fraction and must
    land on the closed-form result."""
    from cmpy_b200 import dmft

    g = np.load(GOLD)
    z = g["z"]
    for u, v_ref in zip(g["u_ed"], g["v_ed"]):
        siam = dmft.twosite_dmft_half_filling(z, float(u), t=1.0, beta=50.0, verbose=False, ref=False, max_iter=100)
        assert abs(siam.v[0] - v_ref) < 1e-7
        gl = dmft.compute_lattice_greens_function(z, siam, 1.0, ref=False)
        assert np.abs(gl - g[f"gf_latt_u{int(u)}"]).max() < 1e-6
    siam = dmft.twosite_dmft_half_filling(z, 4.0, t=1.0, verbose=False, ref=False, max_iter=100)
    assert abs(siam.v[0] - float(g["v_ref"][list(g["u_ref"]).index(4.0)])) < 1e-6
