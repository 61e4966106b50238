"""GPU parity tests: the CUDA path (through the C ABI / ctypes) against the CPU oracle
(oracle/oracle_np.py) on the same inputs, against the committed golden fixtures produced by
the unmodified reference, and through size-independent properties at the BASELINE sizes.

Bars: bit-exact for states / indices / signs / matrix elements; H.v 1e-12 relative;
E0 1e-10; G(z) 1e-8 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
from numpy.testing import assert_array_equal, assert_allclose

import oracle_np as orc

pytestmark = pytest.mark.gpu

HV_RTOL = 1e-12
E0_TOL = 1e-10
GF_TOL = 1e-8


@pytest.fixture(scope="module")
def cm():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import cmpy_b200

    return cmpy_b200


def chain(n, periodic=False):
    return orc.chain_neighbors(n, periodic)


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


# ---------------------------------------------------------------------------------------
# K1: sectors
# ---------------------------------------------------------------------------------------

def test_enumerate_vs_golden(cm, golden):
    from cmpy_b200.basis import enumerate_states

    for L in range(1, 10):
        for n in range(L + 1):
            assert_array_equal(enumerate_states(L, n), golden[f"states_L{L}_n{n}"])
    assert_array_equal(enumerate_states(10, 5), golden["states_L10_n5"])


@pytest.mark.parametrize("L,n", [(12, 6), (16, 8), (20, 10), (24, 3), (31, 2), (40, 2), (62, 1)])
def test_enumerate_vs_oracle(cm, L, n):
    from cmpy_b200.basis import enumerate_states, rank_states

    got = enumerate_states(L, n)
    assert_array_equal(got, orc.enumerate_states(L, n))
    assert got.dtype == np.int64
    assert_array_equal(rank_states(got), np.arange(len(got)))


def test_basis_types_and_order(cm, golden):
    b = cm.Basis(4)
    assert [type(b.get_states(n)).__name__ for n in (None, 0, 1, 2)] == list(golden["states_types"])
    assert list(b.get_states(2)) == [3, 5, 6, 9, 10, 12]
    assert b.get_states(2) is b.get_states(2)  # cached
    sec = b.get_sector(2, 1)
    assert (sec.num_up, sec.num_dn, sec.size) == (6, 4, 24)
    # reference KAT cmpy/tests/test_basis.py:124-141
    for num_sites, n, result in [(2, 1, ["01", "10"]), (3, 2, ["011", "101", "110"]), (3, 3, ["111"])]:
        st = cm.Basis(num_sites).get_states(n)
        assert [cm.binstr(s, num_sites) for s in st] == result
    sb = cm.SpinBasis(6)
    assert sb.get_states(0) == [int(v) for v in orc.enumerate_states(6, 3)]
    with pytest.raises(ValueError):
        cm.SpinBasis(5).get_states(0)


# ---------------------------------------------------------------------------------------
# K2/K3: projectors, exact triplet streams
# ---------------------------------------------------------------------------------------

def _trip(gen):
    r, c, v = [], [], []
    for i, j, val in gen:
        r.append(int(i)); c.append(int(j)); v.append(float(val))
    return np.asarray(r, np.int64), np.asarray(c, np.int64), np.asarray(v, np.float64)


@pytest.mark.parametrize("name,args", [
    ("hop_L4_22_03", (4, 0, 3, 1.0)), ("hop_L4_22_12_t07", (4, 1, 2, 0.7)),
    ("hop_L4_22_03_w0", (0, 0, 3, 1.0)), ("hop_L4_22_03_w2", (2, 0, 3, 1.0)),
])
def test_project_hopping_golden(cm, golden, name, args):
    sec = cm.Basis(4).get_sector(2, 2)
    r, c, v = _trip(cm.project_hopping(sec.up_states, sec.dn_states, *args))
    assert_array_equal(r, golden[name + "_r"])
    assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])


def test_project_hopping_asserts(cm):
    sec = cm.Basis(4).get_sector(2, 2)
    with pytest.raises(AssertionError):
        list(cm.project_hopping(sec.up_states, sec.dn_states, 4, 2, 1, 1.0))


def test_project_diag_golden(cm, golden):
    sec = cm.Basis(4).get_sector(2, 2)
    up, dn = sec.up_states, sec.dn_states
    for name, u in [("inter_L4_22_u4", [4.0] * 4), ("inter_L4_22_uvar", [1.0, 0.0, 2.5, 0.3])]:
        r, c, v = _trip(cm.project_hubbard_inter(up, dn, np.array(u)))
        assert_array_equal(r, golden[name + "_r"]); assert_array_equal(v, golden[name + "_v"])
    for name, eps in [("onsite_L4_22", [0.1, 0.2, 0.3, 0.4]), ("onsite_L4_22_zero", [0, 0, 0.3, 0])]:
        r, c, v = _trip(cm.project_onsite_energy(up, dn, np.array(eps, float)))
        assert_array_equal(r, golden[name + "_r"]); assert_array_equal(c, golden[name + "_c"])
        assert_array_equal(v, golden[name + "_v"])
    sec = cm.Basis(5).get_sector(3, 1)
    r, c, v = _trip(cm.project_hopping(sec.up_states, sec.dn_states, 5, 1, 4, -0.5))
    assert_array_equal(r, golden["hop_L5_31_14_r"]); assert_array_equal(c, golden["hop_L5_31_14_c"])
    assert_array_equal(v, golden["hop_L5_31_14_v"])


def test_projectors_vs_oracle_larger(cm):
    """Signs/targets on periodic and 2-D bonds, incl. non-sector (n=None) string lists."""
    from cmpy_b200.operators import species_hops

    for L, n, bonds in [(8, 4, [(0, 7), (2, 5), (3, 4)]), (9, 3, [(0, 8), (1, 6)]), (12, 6, [(0, 11), (4, 8)])]:
        st = orc.enumerate_states(L, n)
        for (i, j) in bonds:
            for width in (L, 0, j - 1):
                tgt, sgn = species_hops(st, width, i, j)
                o, t, s = orc.species_hops(st, width, i, j)
                assert_array_equal(np.nonzero(tgt >= 0)[0], o)
                assert_array_equal(tgt[o], t)
                assert_array_equal(sgn[o], s)
    st = np.arange(2 ** 6, dtype=np.int64)  # Basis.get_states(None)
    tgt, sgn = species_hops(st, 6, 1, 4)
    o, t, s = orc.species_hops(st, 6, 1, 4)
    assert_array_equal(tgt[o], t); assert_array_equal(sgn[o], s)


HUB = {
    "hub_chain4_22": (4, chain(4), dict(inter=4.0, mu=2.0, hop=1.0), 2, 2),
    "hub_ring4_22": (4, chain(4, True), dict(inter=4.0, mu=2.0, hop=1.0), 2, 2),
    "hub_2x2_22": (4, [[0, 1], [0, 2], [1, 3], [2, 3]], dict(inter=4.0, mu=2.0, hop=1.0), 2, 2),
    "hub_chain5_32": (5, chain(5), dict(inter=3.0, eps=0.25, mu=1.0, hop=-0.8), 3, 2),
    "hub_chain6_33": (6, chain(6), dict(inter=4.0, mu=2.0, hop=1.0), 3, 3),
    "hub_ring6_33": (6, chain(6, True), dict(inter=4.0, mu=2.0, hop=1.0), 3, 3),
    "hub_chain3_10": (3, chain(3), dict(inter=4.0, mu=2.0, hop=1.0), 1, 0),
}


@pytest.mark.parametrize("name", list(HUB))
def test_hubbard_model_golden(cm, golden, name):
    from cmpy_b200.models import HubbardModel

    L, nb, kw, nu, nd = HUB[name]
    model = HubbardModel(L, nb, **kw)
    sec = model.get_sector(nu, nd)
    r, c, v = _trip(model._hamiltonian_data(sec.up_states, sec.dn_states))
    assert_array_equal(r, golden[name + "_r"]); assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])
    hamop = model.hamilton_operator(nu, nd)
    x = np.cos(0.37 * np.arange(hamop.shape[0]))
    ref = golden[name + "_hv"]
    for variant in (1, 2, 3, 5, 6, 7, 0):
        try:
            hamop.set_variant(variant)
            y = hamop.matvec(x)
        except Exception as exc:  # variants 2/3/5-7 need complete sectors whose rows fit shared memory
            assert variant in (2, 3, 5, 6, 7) and "variant" in str(exc), exc
            continue
        assert relerr(y, ref) < HV_RTOL
    hamop.set_variant(0)
    # lazily materialised COO view equals the reference stream
    assert_array_equal(hamop.data, golden[name + "_v"])
    assert_array_equal(hamop.indices[:, 0], golden[name + "_r"])
    # COO-constructed operator (source-compatible constructor)
    data, indices = model.hamiltonian_data(sec.up_states, sec.dn_states)
    coo = cm.HamiltonOperator(hamop.shape[0], data, indices)
    assert relerr(coo.matvec(x), ref) < HV_RTOL
    assert abs(coo.trace() - hamop.trace()) < 1e-10
    dense = hamop.toarray()
    assert_allclose(dense, coo.toarray(), atol=1e-13)
    assert cm.is_hermitian(dense)
    assert abs(np.linalg.eigvalsh(dense)[0] - golden[name + "_e0"]) < 1e-11


def test_hubbard_golden_matrix(cm, golden):
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(2, [[0, 1]], inter=2.0, eps=1.0, hop=1.0)
    expected = [[4.0, 1.0, 1.0, 0.0], [1.0, 2.0, 0.0, 1.0], [1.0, 0.0, 2.0, 1.0], [0.0, 1.0, 1.0, 4.0]]
    assert_array_equal(model.hamiltonian(1, 1), expected)  # cmpy/tests/test_models_hubbard.py:14-25


@pytest.mark.parametrize("name,kw,nu,nd", [
    ("siam4_22", dict(u=2.0, eps_imp=0.0, eps_bath=[0.1, 0.2, 0.3], v=[1.0, 0.7, 0.4]), 2, 2),
    ("siam2_11", dict(u=4.0, v=[1.0], mu=2.0, eps_bath=0.0), 1, 1),
])
def test_siam_golden(cm, golden, name, kw, nu, nd):
    from cmpy_b200.models import SingleImpurityAndersonModel

    model = SingleImpurityAndersonModel(**kw)
    sec = model.get_sector(nu, nd)
    r, c, v = _trip(model._hamiltonian_data(sec.up_states, sec.dn_states))
    assert_array_equal(r, golden[name + "_r"]); assert_array_equal(c, golden[name + "_c"])
    assert_array_equal(v, golden[name + "_v"])
    hamop = model.hamilton_operator(nu, nd)
    x = np.cos(0.37 * np.arange(hamop.shape[0]))
    assert relerr(hamop.matvec(x), golden[name + "_hv"]) < HV_RTOL
    assert abs(np.linalg.eigvalsh(hamop.toarray())[0] - golden[name + "_e0"]) < 1e-11


# ---------------------------------------------------------------------------------------
# K4: H.v
# ---------------------------------------------------------------------------------------

def test_hv_all_sectors_l4(cm, golden):
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(4, chain(4), inter=4.0, mu=2.0, hop=1.0)
    for nu in range(5):
        for nd in range(5):
            h = model.hamilton_operator(nu, nd)
            x = np.cos(0.37 * np.arange(h.shape[0]))
            assert_allclose(h.matvec(x), golden[f"hub_chain4_all_{nu}{nd}_hv"], atol=1e-13)


def test_hv_l8_golden_both_kernels(cm, golden):
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(8, chain(8), inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(4, 4)
    x = np.cos(0.37 * np.arange(4900))
    for variant in (1, 2, 3, 5, 6, 7, 11):
        h.set_variant(variant)
        y = h.matvec(x)
        assert relerr(y, golden["hub_chain8_44_hv"]) < HV_RTOL
    h.set_variant(0)
    assert abs(h.trace() - float(golden["hub_chain8_44_trace"])) < 1e-8
    # complex vectors and (n,1) shapes behave like the reference's _matvec
    yc = h.matvec(x + 2j * x)
    assert relerr(yc, golden["hub_chain8_44_hv"] * (1 + 2j)) < HV_RTOL
    assert h.matvec(x.reshape(-1, 1)).shape == (4900, 1)
    assert h.H is h or h.adjoint() is h


@pytest.mark.parametrize("L,nu,nd,nbfn,kw", [
    (10, 5, 5, lambda: chain(10), dict(inter=4.0, mu=2.0, hop=1.0)),
    (10, 4, 6, lambda: chain(10, True), dict(inter=2.5, eps=0.3, mu=1.0, hop=-0.7)),
    (9, 4, 5, lambda: orc.square_neighbors(3, 3), dict(inter=4.0, mu=2.0, hop=1.0)),
    (12, 6, 6, lambda: chain(12), dict(inter=4.0, mu=2.0, hop=1.0)),
    (12, 2, 9, lambda: orc.square_neighbors(4, 3), dict(inter=1.0, mu=0.2, hop=1.3)),
    (11, 5, 3, lambda: chain(11, True), dict(inter=3.0, eps=-0.4, mu=0.5, hop=0.9)),
    (13, 6, 7, lambda: chain(13), dict(inter=4.0, mu=2.0, hop=1.0)),
    (12, 5, 7, lambda: orc.square_neighbors(3, 4), dict(inter=4.0, mu=2.0, hop=-1.0)),
    (7, 3, 0, lambda: chain(7), dict(inter=4.0, mu=2.0, hop=1.0)),
    (7, 4, 7, lambda: chain(7, True), dict(inter=4.0, mu=2.0, hop=1.0)),
])
def test_hv_vs_oracle(cm, L, nu, nd, nbfn, kw):
    from cmpy_b200.models import HubbardModel

    nb = nbfn()
    model = HubbardModel(L, nb, **kw)
    h = model.hamilton_operator(nu, nd)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(h.shape[0]); x /= np.linalg.norm(x)
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    ref = orc.hubbard_matvec_free(up, dn, nb, kw.get("inter", 0.0), kw.get("eps", 0.0) - kw.get("mu", 0.0),
                                  kw.get("hop", 1.0), x, width=L)
    ran = []
    for variant in (1, 2, 3, 5, 6, 7, 11):
        try:
            h.set_variant(variant)
            y = h.matvec(x)
        except RuntimeError as exc:  # degenerate / odd-length rows have no class-major variant
            assert "variant" in str(exc), exc
            continue
        ran.append(variant)
        assert relerr(y, ref) < HV_RTOL, variant
    from math import comb
    assert 1 in ran and (5 in ran or comb(L, nd) % 2 == 1)
    # accumulate / row-slab entry point used by the sharded operator (dn part only)
    h.set_variant(0)


def test_lanczos_fused_class_major_variants(cm):
    """Fused Lanczos epilogue of the class-major kernel (variants 5-7) against the dense E0."""
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import lanczos_run

    L, nb = 10, orc.square_neighbors(2, 5)
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(5, 5)
    up = orc.enumerate_states(L, 5)
    import scipy.sparse.linalg as sla
    r, c, v = orc.hubbard_triplets(up, up, L, nb, 4.0, -2.0, 1.0)
    import scipy.sparse as sp
    a = sp.csr_matrix((v, (r, c)), shape=(h.shape[0],) * 2)
    e_ref = sla.eigsh(a, k=1, which="SA", tol=0)[0][0]
    for variant in (5, 6, 7, 3, 11):
        h.set_variant(variant)
        res = lanczos_run(h, None, maxit=400, tol=1e-12, resid_tol=1e-9)
        assert abs(res.e0 - e_ref) < 1e-10, (variant, res.e0, e_ref)
    h.set_variant(0)


def test_hv_more_than_32_bonds_tiny_dn_sectors(cm):
    """A lattice with 40 bonds (4 x 5 torus, 20 sites) in sectors whose dn list has <= 64 strings:
    the segment kernel then runs 32-thread CTAs and must still fill all 40 bond slots of its
    per-row tables (round-1 advisor finding, hubbard_seg.cuh)."""
    from cmpy_b200.models import HubbardModel

    nx, ny = 4, 5
    nb = []
    for r in range(ny):
        for c in range(nx):
            i = nx * r + c
            for j in (nx * r + (c + 1) % nx, nx * ((r + 1) % ny) + c):
                nb.append([min(i, j), max(i, j)])
    assert len(nb) == 40
    L = nx * ny
    for kw in (dict(inter=4.0, mu=2.0, hop=1.0), dict(inter=2.0, mu=0.3, hop=-0.8)):
        model = HubbardModel(L, nb, **kw)
        for nu, nd in [(2, 0), (2, 1), (3, 1), (2, 19), (1, 20), (10, 1)]:
            up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
            h = model.hamilton_operator(nu, nd)
            x = np.random.default_rng(3).standard_normal(h.shape[0])
            ref = orc.hubbard_matvec_free(up, dn, nb, kw["inter"], -kw["mu"], kw["hop"], x, width=L)
            for variant in (0, 1, 3):
                try:
                    h.set_variant(variant)
                    y = h.matvec(x)
                except RuntimeError as exc:
                    assert "variant" in str(exc), exc
                    continue
                assert relerr(y, ref) < HV_RTOL, (nu, nd, variant)


def test_hubbard_hamiltonian_csr_helper(cm, golden):
    """`hubbard_hamiltonian` (BASELINE.md section 3 baseline B, cmpy/models/hubbard.py:25-34): the CSR matrix
    of a sector from the projector streams == the dense Hamiltonian of the model == the oracle's triplets."""
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.models.hubbard import hubbard_hamiltonian

    L = 6
    nb = chain(L, True)
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    sec = model.basis.get_sector(3, 2)
    a = hubbard_hamiltonian(sec, nb, inter=4.0, eps=-2.0, hop=1.0)
    n = len(sec.up_states) * len(sec.dn_states)
    assert a.shape == (n, n)
    assert_allclose(a.toarray(), model.hamiltonian(sector=sec), atol=1e-13)
    r, c, v = orc.hubbard_triplets(np.asarray(sec.up_states), np.asarray(sec.dn_states), L, nb, 4.0, -2.0, 1.0)
    assert_allclose(a.toarray(), orc.coo_dense(n, r, c, v), atol=1e-13)


def test_hv_siam_vs_oracle(cm):
    from cmpy_b200.models import SingleImpurityAndersonModel

    kw = dict(u=3.0, eps_imp=-0.2, eps_bath=[0.1, -0.2, 0.3, -0.4, 0.5], v=[1.0, 0.7, 0.4, 0.3, 0.2])
    model = SingleImpurityAndersonModel(**kw)
    L = 6
    for nu, nd in [(3, 3), (2, 4), (0, 3), (6, 1)]:
        up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
        r, c, v = orc.siam_triplets(up, dn, **kw)
        h = model.hamilton_operator(nu, nd)
        x = np.random.default_rng(1).standard_normal(h.shape[0])
        ref = orc.coo_matvec(h.shape[0], r, c, v, x)
        for variant in (1, 0, 3) if min(len(up), len(dn)) > 1 else (1, 0):
            h.set_variant(variant)
            assert relerr(h.matvec(x), ref) < HV_RTOL


@pytest.mark.parametrize("periodic", [False, True])
def test_hamiltonian_hermitian_all_sectors(cm, periodic):
    """cmpy/tests/test_models_hubbard.py:28-58 with stand-in neighbor lists."""
    from cmpy_b200.models import HubbardModel

    for L in (1, 2, 3, 4, 5):
        for u in (0.0, 2.0):
            model = HubbardModel(L, chain(L, periodic), inter=u, mu=u / 2, hop=1.0)
            for nu, nd in model.basis.iter_fillings():
                assert cm.is_hermitian(model.hamiltonian(nu, nd))


def test_hv_torch_zero_copy_and_properties_c4(cm):
    """BASELINE config C4 (4x4 Hubbard, half filling, dim 165 636 900): size-independent
    properties -- symmetry <x,Hy> = <Hx,y>, linearity, agreement of the two kernels, trace."""
    import torch
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(16, orc.square_neighbors(4, 4), inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(8, 8)
    n = h.shape[0]
    assert n == 165636900
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    hx = h.matvec(x)
    hy = h.matvec(y)
    assert hx.is_cuda
    lhs, rhs = float(torch.dot(x, hy)), float(torch.dot(hx, y))
    assert abs(lhs - rhs) < 1e-9 * max(abs(lhs), 1.0) * 10
    hxy = h.matvec(2.0 * x - 0.5 * y)
    assert float((hxy - (2.0 * hx - 0.5 * hy)).abs().max()) < 1e-11 * float(hx.abs().max())
    for variant in (1, 4, 5, 6, 7):
        h.set_variant(variant)
        hx1 = h.matvec(x)
        assert float((hx1 - hx).abs().max()) < 1e-12 * float(hx.abs().max()), variant
    h.set_variant(0)
    d = h.diagonal()
    up = orc.enumerate_states(16, 8)
    e = -2.0 * 16 + 4.0 * np.array([int(up[5] & s).bit_count() for s in up[:100]])
    assert_allclose(d[5 * 12870: 5 * 12870 + 100], e, atol=1e-12)


def _c4_oracle():
    import oracle_c

    return oracle_c.hubbard_oracle(16, 8, 8, orc.square_neighbors(4, 4), 4.0, -2.0, 1.0)


def test_hv_c4_full_vector_vs_oracle(cm):
    """The benchmarked configuration itself (C4, dim 165 636 900, default kernel selection) against
    the C oracle (oracle/hv_oracle.c, pinned to the reference fixtures by tests/test_oracle_c.py):
    the WHOLE result vector, 1e-12 relative; then the fused Lanczos instantiation of the same kernel
    through its first two coefficients (alpha_0 = <v|H|v>, beta_1 = |Hv - alpha_0 v|) computed from the
    oracle's H.v.  ref: cmpy/operators.py:463-527, 626-630."""
    import torch
    from cmpy_b200.exactdiag import lanczos_run
    from cmpy_b200.models import HubbardModel

    oc = _c4_oracle()
    model = HubbardModel(16, orc.square_neighbors(4, 4), inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(8, 8)
    n = h.shape[0]
    assert n == oc.size == 165636900
    x = np.random.default_rng(7).standard_normal(n)
    x /= np.linalg.norm(x)
    ref = oc.matvec(x)
    scale = np.abs(ref).max()
    xd = torch.from_numpy(x).cuda()
    got = h.matvec(xd).cpu().numpy()
    assert np.abs(got - ref).max() / scale < HV_RTOL
    # sample rows of every explicit kernel variant that supports this sector
    rows = [0, 6419, 12838]
    for variant in (1, 4, 5, 11):
        h.set_variant(variant)
        yv = h.matvec(xd)
        for r0 in rows:
            sl = slice(r0 * 12870, (r0 + 32) * 12870)
            assert np.abs(yv[sl].cpu().numpy() - ref[sl]).max() / scale < HV_RTOL, (variant, r0)
        del yv
    h.set_variant(0)
    # fused Lanczos epilogue (the LZ instantiation): two iterations, coefficients from the oracle
    a0 = float(x @ ref)
    r1 = ref - a0 * x
    b1 = float(np.linalg.norm(r1))
    res = lanczos_run(h, xd, maxit=2, tol=0.0, check_every=2)
    assert abs(res.alpha[0] - a0) < 1e-11 * max(1.0, abs(a0))
    assert abs(res.beta[1] - b1) < 1e-11 * max(1.0, abs(b1))
    a1 = float(r1 @ oc.matvec(r1)) / (b1 * b1)
    assert abs(res.alpha[1] - a1) < 1e-10 * max(1.0, abs(a1))


# ---------------------------------------------------------------------------------------
# K6: ladder operators
# ---------------------------------------------------------------------------------------

def test_ladder_up_golden(cm, golden):
    b4 = cm.Basis(4)
    for nu in range(4):
        for nd in range(5):
            s, s1 = b4.get_sector(nu, nd), b4.upper_sector(nu, nd, cm.UP)
            for pos in range(4):
                x = np.cos(0.37 * np.arange(s.size)) + 0.5
                assert_array_equal(cm.CreationOperator(s, s1, pos, cm.UP).matvec(x), golden[f"cdag_L4_{nu}{nd}_p{pos}"])
                x1 = np.cos(0.21 * np.arange(s1.size)) + 0.5
                assert_array_equal(cm.AnnihilationOperator(s1, s, pos, cm.UP).matvec(x1), golden[f"c_L4_{nu + 1}{nd}_p{pos}"])


@pytest.mark.parametrize("num_sites", [2, 3, 4, 5])
@pytest.mark.parametrize("sigma", [1, 2])
def test_creation_annihilation_adjoint(cm, num_sites, sigma):
    """cmpy/tests/test_operator.py:31-42 (the reference itself fails for sigma=DN)."""
    basis = cm.Basis(num_sites)
    for n_up, n_dn in basis.iter_fillings():
        sector = basis.get_sector(n_up, n_dn)
        sector_p1 = basis.upper_sector(n_up, n_dn, sigma)
        if sector_p1 is None:
            continue
        for pos in range(num_sites):
            for signed in (False, True):
                cd = cm.CreationOperator(sector, sector_p1, pos, sigma, signed=signed)
                c = cm.AnnihilationOperator(sector_p1, sector, pos, sigma, signed=signed)
                a, b = cd.toarray(), c.toarray()
                assert a.dtype == np.complex64
                assert_array_equal(b, a.T.conj())
            x = np.random.default_rng(2).standard_normal(sector.size)
            up_t = sector_p1.up_states if sigma == 1 else sector.up_states
            dn_t = sector_p1.dn_states if sigma == 2 else sector.dn_states
            ref = orc.ladder_apply(x, sector.up_states, sector.dn_states, up_t, dn_t, pos, sigma, True)
            assert_array_equal(cm.CreationOperator(sector, sector_p1, pos, sigma).matvec(x), ref)


def test_signed_ladder_anticommutator(cm):
    """{c_i, c^+_j} = delta_ij on the full Fock space of 3 sites (signed mode)."""
    L = 3
    basis = cm.Basis(L)
    full = basis.get_sector(None, None)

    def mat(pos, sigma, dagger):
        # assemble the operator on the full space from its sector blocks
        dim = 4 ** L
        m = np.zeros((dim, dim))
        idx = {(u, d): u * 2 ** L + d for u in range(2 ** L) for d in range(2 ** L)}
        for nu, nd in basis.iter_fillings():
            s = basis.get_sector(nu, nd)
            t = basis.upper_sector(nu, nd, sigma) if dagger else basis.lower_sector(nu, nd, sigma)
            if t is None:
                continue
            cls = cm.CreationOperator if dagger else cm.AnnihilationOperator
            blk = cls(s, t, pos, sigma, signed=True).toarray().real
            up_t = t.up_states if sigma == 1 else s.up_states
            dn_t = t.dn_states if sigma == 2 else s.dn_states
            src = [idx[(int(u), int(d))] for u in s.up_states for d in s.dn_states]
            dst = [idx[(int(u), int(d))] for u in up_t for d in dn_t]
            m[np.ix_(dst, src)] += blk
        return m

    del full
    for (p1, s1), (p2, s2) in [((0, 1), (0, 1)), ((0, 1), (2, 1)), ((1, 1), (1, 2)), ((2, 2), (2, 2)), ((0, 2), (1, 2))]:
        c = mat(p1, s1, False); cd = mat(p2, s2, True)
        anti = c @ cd + cd @ c
        expect = np.eye(4 ** L) if (p1, s1) == (p2, s2) else np.zeros((4 ** L, 4 ** L))
        assert_allclose(anti, expect, atol=1e-14)


# ---------------------------------------------------------------------------------------
# K7: Lanczos / E0
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,e0_ref", [(6, -15.092565319505), (8, -20.235806999130), (10, -25.380618820415)])
def test_lanczos_e0_chain(cm, golden, L, e0_ref):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import lanczos_run

    model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(L // 2, L // 2)
    for use_graph in (True, False):
        res = lanczos_run(h, None, maxit=600, tol=1e-12, resid_tol=1e-9, want_vector=True, use_graph=use_graph)
        assert res.converged
        assert abs(res.e0 - e0_ref) < E0_TOL * 10  # App. B values are printed to 1e-12
        psi = res.vector
        r = h.matvec(psi) - res.e0 * psi
        assert float(r.norm()) < 1e-7
        assert abs(float(psi.norm()) - 1.0) < 1e-10
    if L == 8:
        assert abs(res.e0 - float(golden["hub_chain8_44_e0"])) < E0_TOL
    if L == 6:
        assert abs(res.e0 - float(golden["hub_chain6_33_e0"])) < E0_TOL


def test_lanczos_e0_vs_oracle_misc(cm):
    from cmpy_b200.models import HubbardModel, SingleImpurityAndersonModel
    from cmpy_b200.exactdiag import lanczos_run

    cases = [HubbardModel(9, orc.square_neighbors(3, 3), inter=4.0, mu=2.0, hop=1.0).hamilton_operator(4, 5),
             HubbardModel(8, chain(8, True), inter=6.0, mu=3.0, hop=-1.0).hamilton_operator(3, 4),
             SingleImpurityAndersonModel(u=3.0, eps_bath=[0.1, -0.2, 0.3, -0.4, 0.5],
                                         v=[1.0, 0.7, 0.4, 0.3, 0.2]).hamilton_operator(3, 3)]
    for h in cases:
        e_ref = np.linalg.eigvalsh(h.toarray())[0]
        res = lanczos_run(h, None, maxit=800, tol=1e-12, resid_tol=1e-9)
        assert abs(res.e0 - e_ref) < E0_TOL


def test_compute_groundstate(cm, golden):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import compute_groundstate

    for L in (2, 3, 4):
        model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
        gs = compute_groundstate(model, thresh=10 if L == 4 else 50)
        ref = golden[f"groundstate_chain{L}"]
        assert abs(gs.energy - ref[0]) < E0_TOL
        assert (gs.n_up, gs.n_dn) == (int(ref[1]), int(ref[2]))


def test_u0_e0_analytic_l12(cm):
    """U=0 oracle: E0 = sum of lowest levels of hop*A + (eps-mu)*I per species.
    BASELINE config C2 (L=12 half filling, dim 853 776)."""
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import lanczos_run

    L = 12
    model = HubbardModel(L, chain(L), inter=0.0, mu=0.3, hop=1.0)
    a = np.zeros((L, L))
    for i in range(L - 1):
        a[i, i + 1] = a[i + 1, i] = 1.0
    lev = np.linalg.eigvalsh(a - 0.3 * np.eye(L))
    e_ref = 2 * lev[:6].sum()
    res = lanczos_run(model.hamilton_operator(6, 6), None, maxit=800, tol=1e-12, resid_tol=1e-8)
    assert abs(res.e0 - e_ref) < E0_TOL


def test_reference_lanczos_interface(cm, golden):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200 import exactdiag as ed

    ham = HubbardModel(6, chain(6), inter=4.0, mu=2.0, hop=1.0).hamiltonian(3, 3)
    np.random.seed(1234)
    a, b = ed.lanczos_coeffs(ham, 12)
    assert len(a) == 12 and len(b) == 11
    assert_allclose(a, golden["lanczos_ref_a"], rtol=1e-8)
    assert_allclose(b, golden["lanczos_ref_b"], rtol=1e-8)
    e_gs, vec = ed.lanczos_ground_state(a, b)
    assert abs(e_gs - float(golden["lanczos_ref_egs"])) < 1e-8
    t = ed.lanczos_matrix(a, b)
    assert abs(np.linalg.eigvalsh(t)[0] - e_gs) < 1e-10
    assert_allclose(t @ vec, e_gs * vec, atol=1e-9)


# ---------------------------------------------------------------------------------------
# K8: G(z)
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,pos", [(4, 0), (6, 0), (6, 2), (8, 0)])
def test_gf_continued_fraction_vs_reference_lehmann(cm, golden, L, pos):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import gf_continued_fraction

    model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
    z = golden["z_grid"]
    g, info = gf_continued_fraction(model, z, pos=pos, sigma=cm.UP, num_coeffs=700, return_info=True)
    ref = golden[f"gf0T_chain{L}_p{pos}"]
    assert abs(info["e0"] - float(golden[f"gf0T_chain{L}_p{pos}_e0"])) < E0_TOL
    assert_allclose(info["norms"], golden[f"gf0T_chain{L}_p{pos}_norms"], atol=1e-9)
    assert np.abs(g - ref).max() < GF_TOL


def test_gf_u0_oracle_l10(cm):
    """U=0: G_00(z) of the many-body CF equals gf0_lehmann of the hopping matrix (pos=0)."""
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import gf_continued_fraction
    from cmpy_b200.greens import gf0_lehmann

    L = 10
    model = HubbardModel(L, chain(L), inter=0.0, mu=0.0, hop=1.0)
    z = np.linspace(-6, 6, 1001) + 0.05j
    g = gf_continued_fraction(model, z, pos=0, num_coeffs=400)
    ham0 = np.zeros((L, L))
    for i in range(L - 1):
        ham0[i, i + 1] = ham0[i + 1, i] = 1.0
    g0 = gf0_lehmann(ham0, z=z)[:, 0]
    assert_allclose(g0, orc.gf0_lehmann(ham0, z)[:, 0], atol=1e-12)
    assert np.abs(g - g0).max() < GF_TOL


@pytest.mark.parametrize("L", [2, 3, 4, 5])
def test_gf0_golden(cm, golden, L):
    from cmpy_b200.greens import gf0_lehmann

    ham0 = np.zeros((L, L))
    for i in range(L - 1):
        ham0[i, i + 1] = ham0[i + 1, i] = 1.0
    assert_allclose(gf0_lehmann(ham0, z=golden["z_grid"]), golden[f"gf0_chain{L}"], atol=1e-12)
    with pytest.raises(ValueError):
        gf0_lehmann(ham0, z=golden["z_grid"], mode="nope")


@pytest.mark.parametrize("L,beta", [(2, 10.0), (3, 10.0), (4, 10.0), (4, 50.0)])
def test_gf_lehmann_finite_t_golden(cm, golden, L, beta):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import gf_lehmann

    model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
    d = gf_lehmann(model, golden["z_grid"], beta=beta, pos=0, sigma=cm.UP, occ=True)
    meta = golden[f"gfT_chain{L}_b{int(beta)}_meta"]
    assert np.abs(d.gf - golden[f"gfT_chain{L}_b{int(beta)}"]).max() < GF_TOL
    assert abs(d.gs_energy - meta[0]) < 1e-9 and abs(d.occ - meta[1]) < 1e-9
    assert abs(d.occ_double - meta[2]) < 1e-9


def test_siam_impurity_gf_golden(cm, golden):
    from cmpy_b200.models import SingleImpurityAndersonModel

    siam = SingleImpurityAndersonModel(u=4.0, v=[1.0], mu=2.0, eps_bath=0.0, temp=0.1)
    assert np.abs(siam.impurity_gf(golden["z_grid"]) - golden["gfT_siam2_b10"]).max() < GF_TOL


def test_cf_kernel_vs_oracle(cm):
    from cmpy_b200.exactdiag import cf_eval

    rng = np.random.default_rng(3)
    for m in (1, 2, 7, 511, 512, 513, 1500):
        a = rng.standard_normal(m); b = rng.uniform(0.2, 1.5, size=m - 1)
        z = np.linspace(-4, 4, 333) + 0.07j
        for sign in (+1, -1):
            got = cf_eval(a, b, 0.37, -1.2, z, sign=sign).cpu().numpy()
            if sign > 0:
                ref = orc.cf_eval(a, b, 0.37, z - 1.2)
            else:
                ref = orc.cf_eval(-a, b, 0.37, z + 1.2)
            assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------------------------------
# K5: Heisenberg
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("N", [4, 6, 8, 10])
def test_heisenberg_golden(cm, golden, N):
    from cmpy_b200.models import HeisenbergModel
    from refshim import ChainStandIn

    model = HeisenbergModel(ChainStandIn(N), j=1.0, jz=1.0)
    h = model.hamilton_operator(s=0)
    x = np.cos(0.37 * np.arange(h.shape[0]))
    assert relerr(h.matvec(x), golden[f"heis_chain{N}_s0_hv"]) < HV_RTOL
    assert abs(np.linalg.eigvalsh(h.toarray())[0] - float(golden[f"heis_chain{N}_s0_e0"])) < 1e-11
    if N <= 6:
        r, c, v = _trip(model._hamiltonian_data(model.get_states(0)))
        assert_array_equal(r, golden[f"heis_chain{N}_s0_r"]); assert_array_equal(c, golden[f"heis_chain{N}_s0_c"])
        assert_array_equal(v, golden[f"heis_chain{N}_s0_v"])
        data, indices = model.hamiltonian_data(model.get_states(0))
        coo = cm.HamiltonOperator(h.shape[0], data, indices)
        assert relerr(coo.matvec(x), golden[f"heis_chain{N}_s0_hv"]) < HV_RTOL


def test_heisenberg_misc_golden(cm, golden):
    from cmpy_b200.models import HeisenbergModel
    from refshim import ChainStandIn

    m = HeisenbergModel(ChainStandIn(6, periodic=True), j=0.8, jz=1.3)
    assert_allclose(m.hamiltonian(s=1), golden["heis_ring6_s1_ham"], atol=1e-15)
    for N in (3, 4):
        m = HeisenbergModel(ChainStandIn(N), j=1.0, jz=1.0)
        assert_array_equal(m.hamiltonian(), golden[f"heis_chain{N}_full_ham"])
    m = HeisenbergModel(ChainStandIn(6), j=1.0, jz=0.0)
    assert abs(np.linalg.eigvalsh(m.hamiltonian(s=0))[0] - float(golden["heis_xx6_s0_e0"])) < 1e-12
    with pytest.raises(ValueError):
        HeisenbergModel(ChainStandIn(5)).hamilton_operator(s=0)


@pytest.mark.parametrize("N,s", [(11, 0.5), (12, 0), (13, -1.5), (14, 2), (16, 0)])
def test_heisenberg_vs_oracle(cm, N, s):
    from cmpy_b200.models import HeisenbergModel
    from cmpy_b200.exactdiag import lanczos_run
    from refshim import ChainStandIn

    periodic = N % 2 == 0
    model = HeisenbergModel(ChainStandIn(N, periodic=periodic), j=0.9, jz=1.1)
    h = model.hamilton_operator(s=s)
    st = orc.spin_states(N, s)
    assert h.shape[0] == len(st)
    r, c, v = orc.heisenberg_triplets(st, orc.chain_neighbor_lists(N, periodic), 0.9, 1.1)
    x = np.random.default_rng(4).standard_normal(len(st))
    assert relerr(h.matvec(x), orc.coo_matvec(len(st), r, c, v, x)) < HV_RTOL
    if N <= 12:
        import scipy.sparse as sp
        import scipy.sparse.linalg as sla

        e_ref = sla.eigsh(sp.csr_matrix((v, (r, c)), shape=(len(st),) * 2), k=1, which="SA", tol=0)[0][0]
        res = lanczos_run(h, None, maxit=600, tol=1e-12, resid_tol=1e-9)
        assert abs(res.e0 - e_ref) < E0_TOL


def test_heisenberg_xx_chain_n32_c3(cm):
    """BASELINE config C3 (N=32, Sz=0, dim 601 080 390): XX chain (jz=0) ground-state energy
    equals the free-fermion sum of the N/2 lowest levels of the tridiagonal matrix with
    off-diagonal j/4."""
    from cmpy_b200.models import HeisenbergModel
    from cmpy_b200.exactdiag import lanczos_run
    from refshim import ChainStandIn

    N = 32
    model = HeisenbergModel(ChainStandIn(N), j=1.0, jz=0.0)
    h = model.hamilton_operator(s=0)
    assert h.shape[0] == 601080390
    a = np.zeros((N, N))
    for i in range(N - 1):
        a[i, i + 1] = a[i + 1, i] = 0.25
    e_ref = np.linalg.eigvalsh(a)[: N // 2].sum()
    res = lanczos_run(h, None, maxit=400, tol=1e-11, resid_tol=0.0, check_every=20)
    assert abs(res.e0 - e_ref) < E0_TOL


def test_heisenberg_c3_sample_ranges_vs_oracle(cm):
    """BASELINE config C3 itself (Heisenberg chain N=32, Sz=0, j=jz=1, dim 601 080 390, the long-row spin
    path incl. the odd sub-rows left to the generic kernel): ranges of the result vector at the start, in
    the middle, across a high-word boundary and at the end against the C oracle
    (oracle/hv_oracle.c::orc_heisenberg_hv_range, pinned to the reference fixtures), 1e-12; then the
    fused Lanczos instantiation through alpha_0 = <v|H|v> accumulated from the same ranges is NOT
    possible without the whole H.v, so the Lanczos variant is checked through the symmetric form
    <x|H y> = <H x|y> (ref: cmpy/models/heisenberg.py:19-40)."""
    import torch
    import oracle_c
    from cmpy_b200.models import HeisenbergModel
    from refshim import ChainStandIn

    N = 32
    model = HeisenbergModel(ChainStandIn(N), j=1.0, jz=1.0)
    h = model.hamilton_operator(s=0)
    n = h.shape[0]
    assert n == 601080390
    oc = oracle_c.HeisenbergOracle(N, N // 2, orc.chain_neighbor_lists(N), 1.0, 1.0)
    assert oc.size == n
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = h.matvec(x)
    xh = x.cpu().numpy()
    scale = float(y.abs().max())
    cnt = 200000
    for i0 in (0, n // 3, n // 2 - cnt // 2, n - cnt):
        ref = oc.matvec_range(xh, i0, cnt)
        got = y[i0:i0 + cnt].cpu().numpy()
        assert np.abs(got - ref).max() / scale < HV_RTOL, i0
    del xh
    z = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    hz = h.matvec(z)
    lhs, rhs = float(torch.dot(x, hz)), float(torch.dot(y, z))
    assert abs(lhs - rhs) < 1e-9 * max(abs(lhs), 1.0) * 10


# ---------------------------------------------------------------------------------------
# K9: sharded H.v building blocks (single process; the exchange itself is covered by the
# gloo tests on CPU and by bench.py --gpus N)
# ---------------------------------------------------------------------------------------

def test_sharded_operator_world1_matches_single(cm):
    import torch
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.dist import ShardedHubbardOperator

    for L, nb, nu, nd in [(10, chain(10, True), 5, 5), (9, orc.square_neighbors(3, 3), 4, 3), (12, chain(12), 6, 6)]:
        model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
        h = model.hamilton_operator(nu, nd)
        sh = ShardedHubbardOperator(model, nu, nd)
        x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda")
        ref = h.apply(x)
        got = sh.apply_local(x)
        assert float((got - ref).abs().max()) < 1e-12 * float(ref.abs().max())
        yh = sh.matvec(x.cpu().pin_memory())
        assert not yh.is_cuda and float((yh - ref.cpu()).abs().max()) < 1e-12 * float(ref.abs().max())
        # pipelined host batch of local slabs (the e2e call of bench.py at every N)
        xs = [x.cpu().pin_memory(), (2.0 * x).cpu().pin_memory(), (-x).cpu().pin_memory()]
        outs = [torch.empty(h.shape[0], dtype=torch.float64).pin_memory() for _ in range(3)]
        res = sh.matvec_batch(xs, outs)
        for f, r in zip((1.0, 2.0, -1.0), res):
            assert float((r - f * ref.cpu()).abs().max()) < 1e-12 * float(ref.abs().max())


def test_transpose_copy2d_kernels(cm):
    import torch
    from cmpy_b200 import _lib

    a = torch.randn(70, 131, dtype=torch.float64, device="cuda")
    out = torch.zeros(50, 90, dtype=torch.float64, device="cuda")
    # block rows 3..63, cols 11..58 of `a`, transposed into out[:48, 5:65]
    _lib.check(_lib.lib().cmpy_transpose(_lib.c_void_p(a.data_ptr() + 8 * (3 * 131 + 11)), 60, 47, 131,
                                         _lib.c_void_p(out.data_ptr() + 8 * 5), 90, 0, _lib.stream_ptr()))
    assert torch.equal(out[:47, 5:65], a[3:63, 11:58].t())
    acc = torch.ones(60, 47, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().cmpy_copy2d(_lib.c_void_p(a.data_ptr() + 8 * (3 * 131 + 11)), 60, 47, 131,
                                      _lib.ptr(acc), 47, 1, _lib.stream_ptr()))
    assert torch.equal(acc, a[3:63, 11:58] + 1.0)


def test_host_tensor_matvec(cm):
    import torch
    from cmpy_b200.models import HubbardModel

    h = HubbardModel(8, chain(8), inter=4.0, mu=2.0, hop=1.0).hamilton_operator(4, 4)
    x = torch.randn(4900, dtype=torch.float64).pin_memory()
    y = h.matvec(x)
    assert not y.is_cuda
    assert float((y - h.matvec(x.cuda()).cpu()).abs().max()) == 0.0


def test_numpy_matvec_chunked_roundtrip(cm):
    """The call scipy makes -- matvec(np.ndarray) -- on a vector long enough for the chunk-pipelined host
    round trip (8 pieces): same bits as the device call, fresh results, repeated calls, ragged views."""
    import torch
    from cmpy_b200.models import HubbardModel

    h = HubbardModel(14, chain(14), inter=4.0, mu=2.0, hop=1.0).hamilton_operator(7, 7)
    n = h.shape[0]
    assert n == 3432 * 3432
    rng = np.random.default_rng(2)
    work = rng.standard_normal(3 * n + 5)
    x = work[3:3 + n]                       # a view into a larger work array, like ARPACK's workd
    ref = h.apply(torch.from_numpy(np.ascontiguousarray(x)).cuda()).cpu().numpy()
    y1 = h.matvec(x)
    assert isinstance(y1, np.ndarray) and y1.shape == (n,) and np.array_equal(y1, ref)
    y2 = h.matvec(2.0 * x)
    assert np.array_equal(y1, ref) and y2 is not y1      # the first result was not overwritten
    assert np.array_equal(y2, 2.0 * ref)
    y3 = h.matvec(x.reshape(n, 1))
    assert y3.shape == (n, 1) and np.array_equal(y3[:, 0], ref)
    ys = h.matvec(work[5:5 + 2 * n:2])      # strided view
    rs = h.apply(torch.from_numpy(np.ascontiguousarray(work[5:5 + 2 * n:2])).cuda()).cpu().numpy()
    assert np.array_equal(ys, rs)


# ---------------------------------------------------------------------------------------
# f-2: real-time Green's functions against the unmodified reference
# (tests/golden/reference_tevo.npz, generated by oracle/make_golden_tevo.py)
# ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("L,pos", [(4, 0), (4, 2), (5, 0), (6, 0)])
def test_gf_greater_lesser_golden(cm, L, pos):
    from cmpy_b200.models import HubbardModel
    from cmpy_b200 import exactdiag as ed
    from cmpy_b200.matrix import EigenState

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_tevo.npz"))
    key = f"L{L}_p{pos}"
    model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
    n_up, n_dn = (int(v) for v in g[key + "_gs_sector"])
    gs = EigenState(float(g[key + "_gs_energy"]), g[key + "_gs_state"], n_up, n_dn)
    times = g[key + "_times"]
    t, gg = ed.gf_greater(model, gs, times[0], times[-1], len(times), pos, cm.UP)
    assert_allclose(t, times, atol=1e-14)
    assert np.abs(gg - g[key + "_greater"]).max() < 1e-8      # G(t) to 1e-8, as for G(omega)
    t, gl = ed.gf_lesser(model, gs, times[0], times[-1], len(times), pos, cm.UP)
    assert np.abs(gl - g[key + "_lesser"]).max() < 1e-8
    if L != 5:  # unique ground state: the sweep over sectors finds the same state
        t, gt = ed.gf_tevo(model, times[0], times[-1], len(times), pos, cm.UP)
        assert np.abs(gt - (g[key + "_greater"] - g[key + "_lesser"])).max() < 1e-7


def test_fourier_t2z_analytic(cm):
    """G(t) = -i exp(-i eps t)  ->  G(z) = 1/(z - eps) for Im z > 0 (piecewise-linear rule)."""
    from cmpy_b200 import exactdiag as ed

    eps = 0.7
    t = np.linspace(0.0, 200.0, 40001)
    gt = -1j * np.exp(-1j * eps * t)
    om = np.linspace(-3, 3, 61)
    z, gz = ed.fourier_t2z(t, gt, om, delta=1e-8)
    ref = 1.0 / (z - eps)
    assert np.abs(gz - ref).max() < 2e-5 * np.abs(ref).max()


def test_fourier_t2z_two_poles_exact_piecewise_linear(cm):
    """`fourier_t2z` pinned on a second analytic case.  (i) On a coarse, NON-uniform grid the result must
    equal the exact integral of the piecewise-linear interpolant, written here in an independent algebraic
    form (integration by parts, telescoped boundary terms) -- 1e-12; (ii) on a fine grid it must approach
    the two-pole function a/(z - e1) + b/(z - e2).  gftool itself (cmpy/exactdiag.py:311-316) is absent."""
    from cmpy_b200 import exactdiag as ed

    a, b, e1, e2 = 0.3, 0.7, 0.7, -1.1
    g = lambda t: -1j * (a * np.exp(-1j * e1 * t) + b * np.exp(-1j * e2 * t))
    rng = np.random.default_rng(2)
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.05, 0.4, 300))])
    gt = g(t)
    om = np.linspace(-2.5, 2.5, 41)
    z, gz = ed.fourier_t2z(t, gt, om, eta=0.35)
    iz = 1j * z[:, None]
    e = np.exp(iz * t[None, :])
    slope = (gt[1:] - gt[:-1]) / (t[1:] - t[:-1])
    exact = (e[:, -1] * gt[-1] - e[:, 0] * gt[0]) / iz[:, 0] + (slope[None, :] * (e[:, 1:] - e[:, :-1])).sum(axis=1) / z ** 2
    assert np.abs(gz - exact).max() < 1e-12 * max(1.0, np.abs(exact).max())
    tf = np.linspace(0.0, 150.0, 60001)
    z, gz = ed.fourier_t2z(tf, g(tf), om, delta=1e-9)
    ref = a / (z - e1) + b / (z - e2)
    assert np.abs(gz - ref).max() < 5e-5 * np.abs(ref).max()


def test_expm_multiply_consumer_of_the_gpu_operator(cm):
    """The reference's real-time path hands the Hamiltonian operator to `expm_multiply`
    (cmpy/exactdiag.py:248-273 -> cmpy/linalg/expm_multiply.py:215-299, an adaptation of scipy's, which needs
    `A.trace()`, scalar `*`, `A - mu I`, `A.H` and complex mat-vecs / mat-mats).  Here scipy's own
    `expm_multiply` drives `SectorHamiltonOperator` (matrix-free, on the GPU) and must reproduce
    G^>(t) from this package's spectral-measure evaluation (`gf_greater`) to 1e-8."""
    import scipy.sparse.linalg as sla
    from cmpy_b200 import exactdiag as ed
    from cmpy_b200.matrix import EigenState
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.operators import CreationOperator

    L = 6
    model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
    sec = model.basis.get_sector(3, 3)
    evals, evecs = np.linalg.eigh(model.hamiltonian(sector=sec))
    gs = EigenState(float(evals[0]), evecs[:, 0].copy(), 3, 3)
    start, stop, num = 0.0, 4.0, 21
    times, gg = ed.gf_greater(model, gs, start, stop, num, 0, cm.UP)
    sec_p1 = model.basis.upper_sector(3, 3, cm.UP)
    phi = np.asarray(CreationOperator(sec, sec_p1, pos=0, sigma=cm.UP).matvec(gs.state), dtype=np.complex128)
    hamop = model.hamilton_operator(sector=sec_p1)
    assert abs(hamop.trace() - np.trace(model.hamiltonian(sector=sec_p1))) < 1e-9
    a_op = -1j * hamop                       # scalar * operator keeps a LinearOperator with .H and matmat
    tr = -1j * hamop.trace()
    psi_t = sla.expm_multiply(a_op, phi, start=start, stop=stop, num=num, endpoint=True, traceA=tr)
    ref = -1j * np.exp(1j * gs.energy * times) * (psi_t @ phi.conj())
    assert np.abs(gg - ref).max() < 1e-8


def test_matvec_batch_pipelined_host_vectors(cm):
    """Pipelined host batch (three streams, double buffering) == one blocking matvec per vector."""
    import torch
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(12, chain(12), inter=4.0, mu=2.0, hop=1.0)
    h = model.hamilton_operator(6, 6)
    n = h.shape[0]
    rng = np.random.default_rng(3)
    xs = [torch.from_numpy(rng.standard_normal(n)).pin_memory() for _ in range(5)]
    outs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(5)]
    res = h.matvec_batch(xs, outs)
    assert len(res) == 5
    for x, y in zip(xs, res):
        ref = h.matvec(x.numpy())
        assert relerr(y.numpy(), ref) < HV_RTOL
    # default output buffers: the last two results stay valid
    res = h.matvec_batch(xs)
    assert relerr(res[-1].numpy(), h.matvec(xs[-1].numpy())) < HV_RTOL
    assert relerr(res[-2].numpy(), h.matvec(xs[-2].numpy())) < HV_RTOL
    with pytest.raises(TypeError):
        h.matvec_batch([torch.zeros(3, dtype=torch.float64)])


# ---------------------------------------------------------------------------------------
# K4 long rows (more than 16 sites): sub-row launches of the class-major kernel
# ---------------------------------------------------------------------------------------

def _lattice_3x6():
    nb = []
    for r in range(3):
        for c in range(6):
            i = 6 * r + c
            if c + 1 < 6:
                nb.append([i, i + 1])
            if r + 1 < 3:
                nb.append([i, i + 6])
    return nb


@pytest.mark.parametrize("L,nu,nd,nbfn,kw", [
    (17, 2, 8, lambda: chain(17), dict(inter=4.0, mu=2.0, hop=1.0)),
    (17, 1, 5, lambda: chain(17, True), dict(inter=3.0, eps=0.2, mu=0.5, hop=-0.7)),
    (18, 1, 9, _lattice_3x6, dict(inter=4.0, mu=2.0, hop=1.0)),
    (20, 1, 10, lambda: chain(20), dict(inter=4.0, mu=2.0, hop=1.0)),
    (19, 1, 9, lambda: chain(19) + [[3, 17], [12, 18], [16, 18]], dict(inter=2.0, mu=1.0, hop=0.5)),
])
def test_hv_long_rows(cm, L, nu, nd, nbfn, kw):
    """dn strings of more than 16 sites: the long-row variant (8) against the global-gather kernel
    (1) on the row-slab entry point, and the full H.v assembled from two long-row passes (the
    sharded operator with one rank, all-to-all choreography) against the oracle."""
    import torch
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.dist import ShardedHubbardOperator

    nb = nbfn()
    model = HubbardModel(L, nb, **kw)
    h = model.hamilton_operator(nu, nd)
    n = h.shape[0]
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    xt = torch.from_numpy(x).cuda()
    num_up = len(h.up_states)
    h.set_variant(1)
    ref_dn = h.apply_rows(xt, 0, num_up).cpu().numpy()
    h.set_variant(8)
    got = h.apply_rows(xt, 0, num_up).cpu().numpy()
    assert relerr(got, ref_dn) < HV_RTOL
    acc = torch.ones(n, dtype=torch.float64, device="cuda")
    h.apply_rows(xt, 0, num_up, out=acc, accumulate=True)
    assert relerr(acc.cpu().numpy() - 1.0, ref_dn) < 1e-11
    # a slab that does not start at row 0
    if num_up > 1:
        nd_ = len(h.dn_states)
        part = h.apply_rows(xt[nd_:], 1, num_up - 1).cpu().numpy()
        assert relerr(part, ref_dn[nd_:]) < HV_RTOL
    h.set_variant(0)
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    ref = orc.hubbard_matvec_free(up, dn, nb, kw.get("inter", 0.0), kw.get("eps", 0.0) - kw.get("mu", 0.0),
                                  kw.get("hop", 1.0), x, width=L)
    sh = ShardedHubbardOperator(model, nu, nd, exchange="a2a")
    y = sh.apply_local(xt).cpu().numpy()
    assert relerr(y, ref) < HV_RTOL


# ---------------------------------------------------------------------------------------
# K5 on more than 16 sites: class-major fast path (variant 0 / 5) vs the generic row kernel (1)
# ---------------------------------------------------------------------------------------

class _LadderStandIn:
    """2 x (N/2) ladder: site = 2*rung + leg."""

    def __init__(self, num_sites):
        self.num_sites = num_sites

    def neighbors(self, i):
        n = self.num_sites
        out = [i ^ 1]
        if i - 2 >= 0:
            out.append(i - 2)
        if i + 2 < n:
            out.append(i + 2)
        return out


@pytest.mark.parametrize("N,s,latt,j,jz", [
    (17, 0.5, "chain", 1.0, 1.0), (18, 0, "ring", 0.9, 1.1), (18, 1, "ladder", 1.0, 0.7),
    (20, 0, "chain", 1.0, 1.0), (19, -1.5, "ring", 0.8, 1.3), (22, 3, "chain", 1.0, 0.5),
])
def test_heisenberg_fast_path(cm, N, s, latt, j, jz):
    import torch
    from cmpy_b200.models import HeisenbergModel
    from cmpy_b200.exactdiag import lanczos_run
    from refshim import ChainStandIn

    lat = _LadderStandIn(N) if latt == "ladder" else ChainStandIn(N, periodic=(latt == "ring"))
    model = HeisenbergModel(lat, j=j, jz=jz)
    h = model.hamilton_operator(s=s)
    n = h.shape[0]
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    h.set_variant(1)
    ref = h.matvec(x)
    d_ref = h.diagonal()
    h.set_variant(5)            # raises if the fast path is not available
    y = h.matvec(x)
    assert float((y - ref).abs().max() / ref.abs().max()) < HV_RTOL
    if n <= 200000:
        st = orc.spin_states(N, s)
        nbl = [list(lat.neighbors(i)) for i in range(N)]
        r, c, v = orc.heisenberg_triplets(st, nbl, j, jz)
        xo = x.cpu().numpy()
        assert relerr(y.cpu().numpy(), orc.coo_matvec(len(st), r, c, v, xo)) < HV_RTOL
    # fused Lanczos through the multi-launch operator
    e = []
    for variant in (1, 0):
        h.set_variant(variant)
        res = lanczos_run(h, None, maxit=500, tol=1e-12, resid_tol=1e-9)
        e.append(res.e0)
    assert abs(e[0] - e[1]) < E0_TOL
    assert_allclose(h.diagonal(), d_ref, atol=1e-13)


def test_gf_offdiagonal_continued_fraction(cm):
    """f-4: G_ij(z) by polarisation.  U=0: exactly [(z - h)^-1]_ij of the hopping matrix (every
    hop sign matters); U=4: the dense Lehmann sum built from this package's own dense H and
    signed ladder matrices."""
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import gf_continued_fraction

    z = np.linspace(-5, 5, 301) + 0.1j
    L = 6
    nb = chain(L) + [[0, 5], [1, 4]]
    model = HubbardModel(L, nb, inter=0.0, mu=0.0, hop=1.0)
    h = np.zeros((L, L))
    for a, b in nb:
        h[a, b] = h[b, a] = 1.0
    ginv = np.linalg.inv(z[:, None, None] * np.eye(L)[None] - h[None])
    for i, j in [(0, 3), (2, 5), (1, 1)]:
        g = gf_continued_fraction(model, z, pos=(i, j), n_up=3, n_dn=2, signed=True)
        assert np.abs(g - ginv[:, i, j]).max() < GF_TOL
    # interacting, against dense ED assembled from the same operators
    L = 4
    model = HubbardModel(L, chain(L, True), inter=4.0, mu=2.0, hop=1.0)
    basis = model.basis
    sec, sp1, sm1 = basis.get_sector(2, 2), basis.upper_sector(2, 2, cm.UP), basis.lower_sector(2, 2, cm.UP)
    ev, vec = np.linalg.eigh(model.hamiltonian(sector=sec))
    e0, gs = ev[0], vec[:, 0]
    evp, vp = np.linalg.eigh(model.hamiltonian(sector=sp1))
    evm, vm = np.linalg.eigh(model.hamiltonian(sector=sm1))

    def cd(p):
        return cm.CreationOperator(sec, sp1, p, cm.UP, signed=True).toarray().real

    def c(p):
        return cm.AnnihilationOperator(sec, sm1, p, cm.UP, signed=True).toarray().real

    for i, j in [(0, 1), (0, 2), (1, 3)]:
        ai, aj = vp.T @ (cd(i) @ gs), vp.T @ (cd(j) @ gs)
        bi, bj = vm.T @ (c(i) @ gs), vm.T @ (c(j) @ gs)
        ref = (ai * aj / (z[:, None] - evp[None, :] + e0)).sum(1) + (bi * bj / (z[:, None] + evm[None, :] - e0)).sum(1)
        g = gf_continued_fraction(model, z, pos=(i, j), n_up=2, n_dn=2, gs=(e0, gs), signed=True)
        assert np.abs(g - ref).max() < GF_TOL
