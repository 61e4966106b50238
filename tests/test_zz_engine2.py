"""GPU parity of engine 2 of the class-major H.v kernel (chunked tasks, hoisted per-lane state;
cmpy_b200/csrc/hubbard_cls.cuh, variants 9 / 10 of cmpy_hv_set_variant) against the global-gather
kernel, engine 0 and the CPU oracle.  The same phase bodies are checked lane by lane on the CPU in
tests/test_cls_emulation.py; this file checks the compiled kernels.  (Named zz so that it runs
after the established parity suite.)"""
import numpy as np
import pytest

import oracle_np as orc

pytestmark = pytest.mark.gpu

HV_RTOL = 1e-12


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


@pytest.mark.parametrize("L,nu,nd,nbfn,kw", [
    (8, 4, 4, lambda: chain(8), dict(inter=4.0, mu=2.0, hop=1.0)),
    (10, 5, 5, lambda: chain(10, True), dict(inter=4.0, mu=2.0, hop=1.0)),
    (12, 6, 6, lambda: chain(12), dict(inter=4.0, mu=2.0, hop=1.0)),
    (12, 5, 6, lambda: orc.square_neighbors(4, 3), dict(inter=3.0, eps=0.2, mu=0.5, hop=-0.7)),
    (12, 3, 4, lambda: chain(12), dict(inter=1.5, mu=0.3, hop=0.9)),
])
def test_engine2_vs_oracle(cm, L, nu, nd, nbfn, kw):
    """Full H.v (variant 9) and the dn-only row-slab pass against the oracle and variant 1."""
    import torch
    from cmpy_b200.models import HubbardModel

    nb = nbfn()
    model = HubbardModel(L, nb, **kw)
    h = model.hamilton_operator(nu, nd)
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    rng = np.random.default_rng(L + nu)
    x = rng.standard_normal(h.shape[0])
    ref = orc.hubbard_matvec_free(up, dn, nb, kw.get("inter", 0.0), kw.get("eps", 0.0) - kw.get("mu", 0.0),
                                  kw.get("hop", 1.0), x, width=L)
    try:
        h.set_variant(9)
        y = h.matvec(x)
    except RuntimeError as exc:
        h.set_variant(0)
        pytest.skip(f"engine 2 not available for this sector: {exc}")
    assert relerr(y, ref) < HV_RTOL
    xt = torch.from_numpy(x).cuda()
    num_up = len(h.up_states)
    got = h.apply_rows(xt, 0, num_up).cpu().numpy()
    h.set_variant(1)
    ref_dn = h.apply_rows(xt, 0, num_up).cpu().numpy()
    h.set_variant(0)
    assert relerr(got, ref_dn) < HV_RTOL


def test_engine2_c4_properties(cm):
    """BASELINE config C4 (4x4, dim 1.66e8): engine 2 agrees with engine 0 and the segment kernel
    on the full H.v and on the dn-only slab pass (accumulate and offset slabs included)."""
    import torch
    from cmpy_b200.models import HubbardModel

    nb = orc.square_neighbors(4, 4)
    h = HubbardModel(16, nb, inter=4.0, mu=2.0, hop=1.0).hamilton_operator(8, 8)
    n = h.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    h.set_variant(0)
    ref = h.apply(x)
    h.set_variant(9)
    got = h.apply(x)
    assert float((got - ref).abs().max()) < HV_RTOL * float(ref.abs().max())
    num_up, nd_ = len(h.up_states), len(h.dn_states)
    h.set_variant(5)
    ref_dn = h.apply_rows(x, 0, num_up)
    h.set_variant(9)
    got_dn = h.apply_rows(x, 0, num_up)
    assert float((got_dn - ref_dn).abs().max()) < HV_RTOL * float(ref_dn.abs().max())
    acc = torch.ones(n, dtype=torch.float64, device="cuda")
    h.apply_rows(x, 0, num_up, out=acc, accumulate=True)
    assert float((acc - 1.0 - ref_dn).abs().max()) < 1e-11 * float(ref_dn.abs().max())
    part = h.apply_rows(x[7 * nd_:], 7, 300)
    assert float((part - ref_dn[7 * nd_: 307 * nd_]).abs().max()) < HV_RTOL * float(ref_dn.abs().max())
    h.set_variant(0)


def test_engine2_fused_lanczos(cm):
    """The fused Lanczos epilogue of the engine-2 kernel reproduces E0 of the 12-site chain."""
    from cmpy_b200.exactdiag import lanczos_run
    from cmpy_b200.models import HubbardModel

    h = HubbardModel(12, chain(12), inter=4.0, mu=2.0, hop=1.0).hamilton_operator(6, 6)
    h.set_variant(0)
    ref = lanczos_run(h, None, maxit=300, tol=1e-12)
    h.set_variant(9)
    res = lanczos_run(h, None, maxit=300, tol=1e-12)
    h.set_variant(0)
    assert abs(res.e0 - ref.e0) < 1e-10


@pytest.mark.parametrize("L,nu,nd,nbfn,kw", [
    (17, 2, 8, lambda: chain(17), dict(inter=4.0, mu=2.0, hop=1.0)),
    (18, 1, 9, lambda: [[i, i + 1] for i in range(17)] + [[0, 6], [5, 11], [11, 17]], dict(inter=4.0, mu=2.0, hop=1.0)),
    (20, 1, 10, lambda: chain(20), dict(inter=4.0, mu=2.0, hop=1.0)),
])
def test_engine2_long_rows(cm, L, nu, nd, nbfn, kw):
    """Rows of more than 16 sites: variant 10 (engine 2) against variant 8 (engine 0) and the
    global-gather kernel on the row-slab entry point."""
    import torch
    from cmpy_b200.models import HubbardModel

    model = HubbardModel(L, nbfn(), **kw)
    h = model.hamilton_operator(nu, nd)
    n = h.shape[0]
    x = torch.from_numpy(np.random.default_rng(9).standard_normal(n)).cuda()
    num_up = len(h.up_states)
    h.set_variant(1)
    ref = h.apply_rows(x, 0, num_up)
    try:
        h.set_variant(10)
        got = h.apply_rows(x, 0, num_up)
    except RuntimeError as exc:
        h.set_variant(0)
        pytest.skip(f"engine 2 not available for this sector: {exc}")
    h.set_variant(0)
    assert float((got - ref).abs().max()) < HV_RTOL * float(ref.abs().max())


@pytest.mark.parametrize("N,s,periodic", [(18, 0, True), (20, 0, False), (22, 3, False)])
def test_engine2_heisenberg(cm, N, s, periodic):
    """Heisenberg fast path (sub-row launches of the class-major kernel, spin flavour) with
    engine 2 (variant 9) against the generic row kernel (variant 1) and engine 0 (variant 5)."""
    import torch
    from cmpy_b200.models import HeisenbergModel
    from refshim import ChainStandIn

    h = HeisenbergModel(ChainStandIn(N, periodic=periodic), j=0.9, jz=1.1).hamilton_operator(s=s)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda", generator=g)
    h.set_variant(1)
    ref = h.matvec(x)
    h.set_variant(5)
    y0 = h.matvec(x)
    try:
        h.set_variant(9)
        y2 = h.matvec(x)
    except RuntimeError as exc:
        h.set_variant(0)
        pytest.skip(f"engine 2 not available for this spin sector: {exc}")
    h.set_variant(0)
    assert float((y0 - ref).abs().max() / ref.abs().max()) < HV_RTOL
    assert float((y2 - ref).abs().max() / ref.abs().max()) < HV_RTOL
