"""Pins the CPU oracle (oracle/oracle_np.py) to the reference: every fixture in
tests/golden/reference_golden.npz was produced by the unmodified reference
(oracle/make_golden.py). CPU-only."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal, assert_allclose

import oracle_np as orc


def chain(n, periodic=False):
    return orc.chain_neighbors(n, periodic)


def test_states_all(golden):
    for L in range(1, 10):
        for n in range(L + 1):
            assert_array_equal(orc.enumerate_states(L, n), golden[f"states_L{L}_n{n}"])
    assert_array_equal(orc.enumerate_states(10, 5), golden["states_L10_n5"])


def test_reference_kat_basis_order():
    # cmpy/tests/test_basis.py:124-141
    assert list(orc.enumerate_states(3, 2)) == [0b011, 0b101, 0b110]
    assert list(orc.enumerate_states(2, 1)) == [0b01, 0b10]
    assert list(orc.enumerate_states(4, 2)) == [3, 5, 6, 9, 10, 12]


def test_rank_is_bisect():
    for L, n in [(6, 3), (9, 4), (10, 5), (12, 6)]:
        st = orc.enumerate_states(L, n)
        assert_array_equal([orc.rank_state(s) for s in st], np.arange(len(st)))


@pytest.mark.parametrize("name,args", [
    ("hop_L4_22_03", (4, 0, 3, 1.0)), ("hop_L4_22_12_t07", (4, 1, 2, 0.7)),
    ("hop_L4_22_03_w0", (0, 0, 3, 1.0)), ("hop_L4_22_03_w2", (2, 0, 3, 1.0)),
])
def test_project_hopping(golden, name, args):
    up = dn = orc.enumerate_states(4, 2)
    r, c, v = orc.hopping_triplets(up, dn, *args)
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])


def test_project_hopping_l5(golden):
    r, c, v = orc.hopping_triplets(orc.enumerate_states(5, 3), orc.enumerate_states(5, 1), 5, 1, 4, -0.5)
    assert_array_equal(r, golden["hop_L5_31_14_r"])
    assert_array_equal(c, golden["hop_L5_31_14_c"])
    assert_array_equal(v, golden["hop_L5_31_14_v"])


@pytest.mark.parametrize("name,u", [("inter_L4_22_u4", [4.0] * 4), ("inter_L4_22_uvar", [1.0, 0.0, 2.5, 0.3])])
def test_project_inter(golden, name, u):
    up = dn = orc.enumerate_states(4, 2)
    r, c, v = orc.inter_triplets(up, dn, np.array(u))
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(v, golden[name + "_v"])


@pytest.mark.parametrize("name,eps", [("onsite_L4_22", [0.1, 0.2, 0.3, 0.4]), ("onsite_L4_22_zero", [0, 0, 0.3, 0])])
def test_project_onsite(golden, name, eps):
    up = dn = orc.enumerate_states(4, 2)
    r, c, v = orc.onsite_triplets(up, dn, np.array(eps, float))
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])  # bit-exact incl. 0.30000000000000004


HUB = {
    "hub_chain4_22": (4, chain(4), 4.0, -2.0, 1.0, 2, 2),
    "hub_ring4_22": (4, chain(4, True), 4.0, -2.0, 1.0, 2, 2),
    "hub_2x2_22": (4, [[0, 1], [0, 2], [1, 3], [2, 3]], 4.0, -2.0, 1.0, 2, 2),
    "hub_chain5_32": (5, chain(5), 3.0, 0.25 - 1.0, -0.8, 3, 2),
    "hub_chain6_33": (6, chain(6), 4.0, -2.0, 1.0, 3, 3),
    "hub_ring6_33": (6, chain(6, True), 4.0, -2.0, 1.0, 3, 3),
    "hub_chain3_10": (3, chain(3), 4.0, -2.0, 1.0, 1, 0),
}


@pytest.mark.parametrize("name", list(HUB))
def test_hubbard_stream_hv_e0(golden, name):
    L, nb, inter, eps, hop, nu, nd = HUB[name]
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    r, c, v = orc.hubbard_triplets(up, dn, L, nb, inter, eps, hop)
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])
    size = len(up) * len(dn)
    x = np.cos(0.37 * np.arange(size))
    y = orc.coo_matvec(size, r, c, v, x)
    assert_allclose(y, golden[name + "_hv"], rtol=0, atol=1e-13 * np.abs(golden[name + "_hv"]).max())
    y2 = orc.hubbard_matvec_free(up, dn, nb, inter, eps, hop, x, width=L)
    assert_allclose(y2, golden[name + "_hv"], rtol=0, atol=1e-13 * max(1.0, np.abs(y).max()))
    e0 = np.linalg.eigvalsh(orc.coo_dense(size, r, c, v))[0]
    assert abs(e0 - golden[name + "_e0"]) < 1e-11


def test_hubbard_golden_matrix(golden):
    # cmpy/tests/test_models_hubbard.py:14-25
    up = dn = orc.enumerate_states(2, 1)
    r, c, v = orc.hubbard_triplets(up, dn, 2, [[0, 1]], 2.0, 1.0, 1.0)
    ham = orc.coo_dense(4, r, c, v)
    expected = [[4.0, 1.0, 1.0, 0.0], [1.0, 2.0, 0.0, 1.0], [1.0, 0.0, 2.0, 1.0], [0.0, 1.0, 1.0, 4.0]]
    assert_array_equal(ham, expected)
    assert_array_equal(ham, golden["hub2_11_ham"])


@pytest.mark.parametrize("name,kw,nu,nd,L", [
    ("siam4_22", dict(u=2.0, eps_imp=0.0, eps_bath=[0.1, 0.2, 0.3], v=[1.0, 0.7, 0.4]), 2, 2, 4),
    ("siam2_11", dict(u=4.0, eps_imp=0.0, eps_bath=0.0, v=[1.0], mu=2.0), 1, 1, 2),
])
def test_siam_stream(golden, name, kw, nu, nd, L):
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    r, c, v = orc.siam_triplets(up, dn, **kw)
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])
    assert v[r != c].min() >= 0.0  # signless hops (anderson.py:149)


def test_hv_l8_and_e0(golden):
    up = dn = orc.enumerate_states(8, 4)
    x = np.cos(0.37 * np.arange(4900))
    y = orc.hubbard_matvec_free(up, dn, chain(8), 4.0, -2.0, 1.0, x)
    ref = golden["hub_chain8_44_hv"]
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-13
    # SURVEY App. B spot values
    assert abs(ref[0] - 1.652143581426160e+00) < 1e-12
    assert abs(np.linalg.norm(ref) - 3.722570996119713e+02) < 1e-9
    assert abs(float(golden["hub_chain8_44_e0"]) - (-20.235806999130)) < 1e-9


def test_hv_all_sectors_l4(golden):
    for nu in range(5):
        for nd in range(5):
            up, dn = orc.enumerate_states(4, nu), orc.enumerate_states(4, nd)
            size = len(up) * len(dn)
            r, c, v = orc.hubbard_triplets(up, dn, 4, chain(4), 4.0, -2.0, 1.0)
            x = np.cos(0.37 * np.arange(size))
            assert_allclose(orc.coo_matvec(size, r, c, v, x), golden[f"hub_chain4_all_{nu}{nd}_hv"], atol=1e-13)


@pytest.mark.parametrize("N", [4, 6, 8, 10])
def test_heisenberg(golden, N):
    st = orc.spin_states(N, 0)
    r, c, v = orc.heisenberg_triplets(st, orc.chain_neighbor_lists(N), 1.0, 1.0)
    if N <= 6:
        assert_array_equal(r, golden[f"heis_chain{N}_s0_r"])
        assert_array_equal(c, golden[f"heis_chain{N}_s0_c"])
        assert_array_equal(v, golden[f"heis_chain{N}_s0_v"])
    x = np.cos(0.37 * np.arange(len(st)))
    assert_allclose(orc.coo_matvec(len(st), r, c, v, x), golden[f"heis_chain{N}_s0_hv"], atol=1e-13)
    e0 = np.linalg.eigvalsh(orc.coo_dense(len(st), r, c, v))[0]
    assert abs(e0 - golden[f"heis_chain{N}_s0_e0"]) < 1e-12


def test_heisenberg_misc(golden):
    st = orc.spin_states(6, 1)
    r, c, v = orc.heisenberg_triplets(st, orc.chain_neighbor_lists(6, True), 0.8, 1.3)
    # duplicates are summed by scipy's csr in a different order -> 1 ulp
    assert_allclose(orc.coo_dense(len(st), r, c, v), golden["heis_ring6_s1_ham"], rtol=0, atol=5e-16)
    for N in (3, 4):
        st = orc.spin_states(N, None)
        r, c, v = orc.heisenberg_triplets(st, orc.chain_neighbor_lists(N), 1.0, 1.0)
        assert_array_equal(orc.coo_dense(len(st), r, c, v), golden[f"heis_chain{N}_full_ham"])
    with pytest.raises(ValueError):
        orc.spin_states(5, 0)


def test_heisenberg_kron_oracle():
    # cmpy/tests/test_models_heisenberg.py:15-35 (Kronecker-product XXZ)
    sz = np.array([[1, 0], [0, -1]], float); sp = np.array([[0, 1], [0, 0]], float); sm = sp.T
    for N in (3, 4, 5):
        ham = np.zeros((2 ** N, 2 ** N))
        for i in range(N - 1):
            parts = [np.eye(2)] * N
            def kron(ms):
                out = np.eye(1)
                for m in ms:
                    out = np.kron(out, m)
                return out
            two = 0.5 * 0.5 * (np.kron(sp, sm) + np.kron(sm, sp)) + 0.5 * np.kron(sz, sz)
            ham += kron([np.eye(2)] * i + [two] + [np.eye(2)] * (N - i - 2))
        st = orc.spin_states(N, None)
        r, c, v = orc.heisenberg_triplets(st, orc.chain_neighbor_lists(N), 1.0, 1.0)
        assert_array_equal(orc.coo_dense(len(st), r, c, v), ham)


def test_ladder_up(golden):
    for nu in range(4):
        for nd in range(5):
            up, dn = orc.enumerate_states(4, nu), orc.enumerate_states(4, nd)
            up1 = orc.enumerate_states(4, nu + 1)
            for pos in range(4):
                x = np.cos(0.37 * np.arange(len(up) * len(dn))) + 0.5
                y = orc.ladder_apply(x, up, dn, up1, dn, pos, orc.UP, True)
                assert_array_equal(y, golden[f"cdag_L4_{nu}{nd}_p{pos}"])
                x1 = np.cos(0.21 * np.arange(len(up1) * len(dn))) + 0.5
                y1 = orc.ladder_apply(x1, up1, dn, up, dn, pos, orc.UP, False)
                assert_array_equal(y1, golden[f"c_L4_{nu + 1}{nd}_p{pos}"])


def _sector_solver(L, nb, inter, eps, hop):
    def solve(nu, nd):
        up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
        r, c, v = orc.hubbard_triplets(up, dn, L, nb, inter, eps, hop)
        ev, evec = np.linalg.eigh(orc.coo_dense(len(up) * len(dn), r, c, v))
        return ev, evec, up, dn
    return solve


@pytest.mark.parametrize("L,pos", [(4, 0), (6, 0), (6, 2)])
def test_zero_t_lehmann_and_cf(golden, L, pos):
    n = L // 2
    z = golden["z_grid"]
    solve = _sector_solver(L, chain(L), 4.0, -2.0, 1.0)
    ev, evec, up, dn = solve(n, n)
    e0, gs = ev[0], evec[:, 0]
    evp, evecp, upp, _ = solve(n + 1, n)
    evm, evecm, upm, _ = solve(n - 1, n)
    cd = orc.ladder_apply(gs, up, dn, upp, dn, pos, orc.UP, True)
    c = orc.ladder_apply(gs, up, dn, upm, dn, pos, orc.UP, False)
    G = orc.zero_t_lehmann(z, e0, gs, evp, evecp, cd, evm, evecm, c)
    ref = golden[f"gf0T_chain{L}_p{pos}"]
    assert np.abs(G - ref).max() < 1e-10
    # Lanczos continued fraction reproduces the Lehmann sum (SURVEY 8(c) validity check)
    Hp = evecp @ np.diag(evp) @ evecp.T
    Hm = evecm @ np.diag(evm) @ evecm.T
    a, b, n0 = orc.lanczos_coeffs_normalised(lambda v: Hp @ v, cd, 300)
    Gp = orc.cf_eval(a, b, n0 ** 2, z + e0)
    a, b, n0 = orc.lanczos_coeffs_normalised(lambda v: Hm @ v, c, 300)
    Gm = orc.cf_eval(-a, b, n0 ** 2, z - e0)
    assert np.abs(Gp + Gm - ref).max() < 1e-8


@pytest.mark.parametrize("L,beta", [(2, 10.0), (3, 10.0), (4, 10.0), (4, 50.0)])
def test_gf_lehmann_finite_t(golden, L, beta):
    z = golden["z_grid"]
    gf, part, egs, occ, occ2 = orc.gf_lehmann_finite_t(_sector_solver(L, chain(L), 4.0, -2.0, 1.0), L, z, beta)
    ref = golden[f"gfT_chain{L}_b{int(beta)}"]
    meta = golden[f"gfT_chain{L}_b{int(beta)}_meta"]
    assert np.abs(gf - ref).max() < 1e-10
    assert abs(egs - meta[0]) < 1e-10 and abs(occ - meta[1]) < 1e-10 and abs(occ2 - meta[2]) < 1e-10


@pytest.mark.parametrize("L", [2, 3, 4, 5])
def test_gf0(golden, L):
    ham0 = np.zeros((L, L))
    for i in range(L - 1):
        ham0[i, i + 1] = ham0[i + 1, i] = 1.0
    assert_allclose(orc.gf0_lehmann(ham0, golden["z_grid"]), golden[f"gf0_chain{L}"], atol=1e-12)


def test_reference_lanczos(golden):
    up = dn = orc.enumerate_states(6, 3)
    r, c, v = orc.hubbard_triplets(up, dn, 6, chain(6), 4.0, -2.0, 1.0)
    ham = orc.coo_dense(400, r, c, v)
    a, b = orc.reference_lanczos_coeffs(ham, golden["lanczos_ref_psi0"], 12)
    assert_allclose(a, golden["lanczos_ref_a"], rtol=1e-12)
    assert_allclose(b, golden["lanczos_ref_b"], rtol=1e-12)
    # normalised recurrence gives the same tridiagonal matrix
    a2, b2, _ = orc.lanczos_coeffs_normalised(lambda q: ham @ q, golden["lanczos_ref_psi0"], 12)
    assert_allclose(a2, a, rtol=1e-9)
    assert_allclose(b2, b, rtol=1e-9)


@pytest.mark.parametrize("L,pos", [(4, 0), (4, 2), (5, 0), (6, 0)])
def test_realtime_gf_oracle_vs_reference(L, pos):
    """Oracle restatement of gf_greater / gf_lesser against the unmodified reference's output
    (tests/golden/reference_tevo.npz, oracle/make_golden_tevo.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_tevo.npz"))
    key = f"L{L}_p{pos}"
    nb = orc.chain_neighbors(L)
    n_up, n_dn = (int(v) for v in g[key + "_gs_sector"])
    for greater, name in ((True, "_greater"), (False, "_lesser")):
        out = orc.gf_realtime_dense(L, nb, 4.0, -2.0, 1.0, n_up, n_dn, float(g[key + "_gs_energy"]),
                                    g[key + "_gs_state"], g[key + "_times"], pos, greater)
        assert np.abs(out - g[key + name]).max() < 1e-10


def test_sz_correlator_is_the_diagonal_of_a_one_pair_ising_model():
    """The identity cmpy_b200/observables.py rests on: the diagonal of the reference's Heisenberg
    Hamiltonian with the single directed pair (i, j), j = 0, jz = 1 is Sz_i Sz_j
    (cmpy/models/heisenberg.py:28-31), so scripts/heisenberg.py's sz_correl is a weighted sum of it."""
    N, s = 8, 0
    states = orc.spin_states(N, s)
    rng = np.random.default_rng(2)
    gs = rng.standard_normal(len(states))
    gs /= np.linalg.norm(gs)
    for i, j in [(0, 3), (5, 2), (0, 7)]:
        nbl = [[] for _ in range(N)]
        nbl[i] = [j]
        r, c, v = orc.heisenberg_triplets(states, nbl, j=0.0, jz=1.0)
        diag = np.zeros(len(states))
        np.add.at(diag, r[r == c], v[r == c])
        assert set(np.round(diag, 12)) <= {0.25, -0.25}
        if i == 0:
            assert abs(np.dot(gs * gs, diag) * 1.7 - orc.sz_correl(states, gs, j, j=1.7)) < 1e-14
        else:
            lo, d = min(i, j), abs(i - j)
            assert abs(np.dot(gs * gs, diag) - orc.sz_correl(states, gs, d, pos=lo)) < 1e-14
    # total magnetisation: sum_k <Sz_k> = s
    assert abs(sum(orc.sz_expval(states, gs, k) for k in range(N)) - s) < 1e-13
