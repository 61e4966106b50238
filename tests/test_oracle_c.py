"""Pins the C oracle (oracle/hv_oracle.c) to the numpy oracle and the reference fixtures."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle_np as orc
import oracle_c as orcc


def test_c_enumerate():
    for L, n in [(4, 2), (9, 4), (12, 6), (16, 8), (5, 0), (5, 5)]:
        assert_array_equal(orcc.enumerate_states(L, n), orc.enumerate_states(L, n))


def test_c_hv_golden(golden):
    o = orcc.hubbard_oracle(8, 4, 4, orc.chain_neighbors(8), 4.0, -2.0, 1.0)
    x = np.cos(0.37 * np.arange(4900))
    ref = golden["hub_chain8_44_hv"]
    assert np.abs(o.matvec(x) - ref).max() / np.abs(ref).max() < 1e-13
    assert_array_equal(o.matvec_rows(x, 10, 5, nthreads=1), o.matvec(x)[700:1050])


@pytest.mark.parametrize("L,nu,nd,nb", [(6, 3, 3, orc.chain_neighbors(6, True)), (9, 4, 5, orc.square_neighbors(3, 3)),
                                       (10, 5, 5, orc.chain_neighbors(10))])
def test_c_hv_vs_numpy(L, nu, nd, nb):
    o = orcc.hubbard_oracle(L, nu, nd, nb, 3.0, -0.7, 0.9)
    x = np.random.default_rng(0).standard_normal(o.size)
    ref = orc.hubbard_matvec_free(o.up, o.dn, nb, 3.0, -0.7, 0.9, x, width=L)
    assert np.abs(o.matvec(x) - ref).max() / np.abs(ref).max() < 1e-13


def test_c_siam_signless(golden):
    up = dn = orc.enumerate_states(4, 2)
    o = orcc.HubbardOracle(4, up, dn, [(0, 1), (0, 2), (0, 3)], [1.0, 0.7, 0.4],
                           [0.0 - 1.0, 0.1, 0.2, 0.3], [2.0, 0, 0, 0], 0)
    x = np.cos(0.37 * np.arange(36))
    ref = golden["siam4_22_hv"]
    assert np.abs(o.matvec(x) - ref).max() < 1e-13


@pytest.mark.parametrize("N", [4, 6, 8, 10])
def test_c_heisenberg_golden(golden, N):
    """C restatement of the Heisenberg H.v against the H.v produced by the unmodified reference."""
    o = orcc.HeisenbergOracle(N, N // 2, orc.chain_neighbor_lists(N), 1.0, 1.0)
    x = np.cos(0.37 * np.arange(o.size))
    ref = golden[f"heis_chain{N}_s0_hv"]
    assert np.abs(o.matvec(x) - ref).max() <= 1e-13 * np.abs(ref).max()
    assert_array_equal(o.matvec_range(x, 3, 2, nthreads=1), o.matvec(x)[3:5])


@pytest.mark.parametrize("N,n_up,periodic,j,jz", [(12, 6, False, 0.9, 1.1), (11, 4, True, 1.0, -0.4), (14, 7, True, 1.3, 0.0)])
def test_c_heisenberg_vs_numpy(N, n_up, periodic, j, jz):
    nbl = orc.chain_neighbor_lists(N, periodic)
    st = orc.enumerate_states(N, n_up)
    r, c, v = orc.heisenberg_triplets(st, nbl, j, jz)
    x = np.random.default_rng(1).standard_normal(len(st))
    ref = orc.coo_matvec(len(st), r, c, v, x)
    o = orcc.HeisenbergOracle(N, n_up, nbl, j, jz)
    assert o.size == len(st)
    assert np.abs(o.matvec(x) - ref).max() <= 1e-13 * np.abs(ref).max()


def test_c_hv_random_graphs():
    """Random bond lists (any pair i < j, also long-range), random fillings and both sign conventions
    (width = L and the Anderson width = 0): C oracle == numpy oracle == dense matrix of the triplet oracle."""
    rng = np.random.default_rng(20260)
    for trial in range(12):
        L = int(rng.integers(3, 9))
        nb = sorted({tuple(sorted(rng.choice(L, size=2, replace=False).tolist())) for _ in range(int(rng.integers(1, 2 * L)))})
        nb = [(int(i), int(j)) for i, j in nb]
        nu, nd = int(rng.integers(0, L + 1)), int(rng.integers(0, L + 1))
        width = L if trial % 3 else 0
        inter, eps, hop = float(rng.normal()), float(rng.normal()), float(rng.normal())
        o = orcc.hubbard_oracle(L, nu, nd, nb, inter, eps, hop, width=width)
        x = rng.standard_normal(o.size)
        ref = orc.hubbard_matvec_free(o.up, o.dn, nb, inter, eps, hop, x, width=width)
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(o.matvec(x) - ref).max() <= 1e-13 * scale, (L, nb, nu, nd, width)
        if width == L and 0 < o.size <= 2000:      # the reference's own triplet stream (signs up to num_sites)
            r, c, v = orc.hubbard_triplets(o.up, o.dn, L, nb, inter, eps, hop)
            y = orc.coo_matvec(o.size, np.asarray(r), np.asarray(c), np.asarray(v, dtype=np.float64), x)
            assert np.abs(y - ref).max() <= 1e-12 * scale
