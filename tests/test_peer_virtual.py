"""K9 on ONE GPU with 'virtual ranks': the peer-transpose kernels of the sharded H.v
(`cmpy_transpose_push`, `cmpy_transpose_pull_acc`, the default data plane of
cmpy_b200/dist.py::_apply_local_peer) take a table of slab pointers; here every 'peer' slab is a
local buffer, so the exact kernels the N-GPU step launches run under the single-GPU driver suite.

  * push / pull against their definition (include/cmpy_b200.h), ragged slab bounds
  * the whole sharded sequence (dn pass -> push -> up pass in the dn-major slab -> pull-accumulate)
    over W virtual ranks == the single-GPU H.v == the CPU oracle (1e-12)
Layout being sharded: cmpy/operators.py:33-90 (idx = up_idx * num_dn + dn_idx)."""
import ctypes
import os

import numpy as np
import pytest

import oracle_np as orc

pytestmark = pytest.mark.gpu


def _bounds(n, world):
    return [(n * k) // world for k in range(world + 1)]


@pytest.mark.parametrize("world,nu,nd", [(2, 70, 131), (3, 70, 131), (5, 33, 64), (8, 257, 40),
                                         # odd slab starts, several row tiles, the 4x4 row count
                                         (4, 1198, 97), (7, 1000, 33), (8, 12870, 45)])
def test_push_and_pull_definitions(world, nu, nd):
    import torch
    from cmpy_b200 import _lib

    L = _lib.lib()
    rb, cb = _bounds(nu, world), _bounds(nd, world)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(nu, nd, dtype=torch.float64, device="cuda", generator=g)
    xts = [torch.full(((cb[q + 1] - cb[q]) * nu,), float("nan"), dtype=torch.float64, device="cuda")
           for q in range(world)]
    peers = (ctypes.c_void_p * world)(*[t.data_ptr() for t in xts])
    bounds = (ctypes.c_int64 * (world + 1))(*cb)
    for p in range(world):          # every virtual rank pushes its slab of up-rows
        slab = x[rb[p]:rb[p + 1]].contiguous()
        _lib.check(L.cmpy_transpose_push(_lib.ptr(slab), rb[p + 1] - rb[p], nd, rb[p], nu, world, bounds,
                                         peers, _lib.stream_ptr()))
    torch.cuda.synchronize()
    for q in range(world):          # XT_q = X[:, cols of q]^T, bit for bit
        assert torch.equal(xts[q].view(cb[q + 1] - cb[q], nu), x[:, cb[q]:cb[q + 1]].t())
    # pull: y[r, c] += YT_q[c - cb[q], row0 + r]
    yts = [torch.randn_like(t) for t in xts]
    peers_y = (ctypes.c_void_p * world)(*[t.data_ptr() for t in yts])
    for p in range(world):
        nrows = rb[p + 1] - rb[p]
        y0 = torch.randn(nrows, nd, dtype=torch.float64, device="cuda", generator=g)
        y = y0.clone()
        _lib.check(L.cmpy_transpose_pull_acc(_lib.ptr(y), nrows, nd, rb[p], nu, world, bounds, peers_y,
                                             _lib.stream_ptr()))
        torch.cuda.synchronize()
        ref = y0.clone()
        for q in range(world):
            ref[:, cb[q]:cb[q + 1]] += yts[q].view(cb[q + 1] - cb[q], nu)[:, rb[p]:rb[p + 1]].t()
        assert torch.equal(y, ref)


@pytest.mark.parametrize("L,nu_f,nd_f,nbfn,world", [
    (10, 5, 5, lambda: orc.chain_neighbors(10), 2),
    (10, 4, 6, lambda: orc.chain_neighbors(10, True), 3),
    (12, 6, 6, lambda: orc.square_neighbors(4, 3), 4),
    (12, 5, 7, lambda: orc.chain_neighbors(12), 8),
])
def test_sharded_sequence_with_virtual_ranks(L, nu_f, nd_f, nbfn, world):
    """dn pass + push + up pass + pull-accumulate, exactly the calls of _apply_local_peer /
    _apply_second_half, for W virtual ranks on one device."""
    import torch
    from cmpy_b200 import _lib
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.operators import SectorHamiltonOperator

    lib = _lib.lib()
    nb = nbfn()
    kw = dict(inter=4.0, mu=2.0, hop=1.0)
    model = HubbardModel(L, nb, **kw)
    spec = model._operator_spec()
    up, dn = orc.enumerate_states(L, nu_f), orc.enumerate_states(L, nd_f)
    nu, nd = len(up), len(dn)
    op_main = SectorHamiltonOperator(L, up, dn, spec["bonds"], spec["hops"], spec["eps"], spec["u"],
                                     spec["sign_width"])
    zeros = np.zeros(L)
    op_t = SectorHamiltonOperator(L, dn, up, spec["bonds"], spec["hops"], zeros, zeros, spec["sign_width"])
    rb, cb = _bounds(nu, world), _bounds(nd, world)
    xh = np.random.default_rng(5).standard_normal(nu * nd)
    x = torch.from_numpy(xh).cuda()
    xts = [torch.zeros(max((cb[q + 1] - cb[q]) * nu, 1), dtype=torch.float64, device="cuda") for q in range(world)]
    yts = [torch.zeros_like(t) for t in xts]
    peers_x = (ctypes.c_void_p * world)(*[t.data_ptr() for t in xts])
    peers_y = (ctypes.c_void_p * world)(*[t.data_ptr() for t in yts])
    bounds = (ctypes.c_int64 * (world + 1))(*cb)
    y = torch.empty_like(x)
    for p in range(world):      # first half on every rank: local dn pass + push
        r0, r1 = rb[p], rb[p + 1]
        xs = x[r0 * nd:r1 * nd]
        op_main.apply_rows(xs, r0, r1 - r0, out=y[r0 * nd:r1 * nd])
        _lib.check(lib.cmpy_transpose_push(_lib.ptr(xs), r1 - r0, nd, r0, nu, world, bounds, peers_x,
                                           _lib.stream_ptr()))
    for q in range(world):      # up hops, row-local in the dn-major slab of rank q
        if cb[q + 1] > cb[q]:
            op_t.apply_rows(xts[q], cb[q], cb[q + 1] - cb[q], out=yts[q])
    for p in range(world):      # second half: pull-accumulate
        r0, r1 = rb[p], rb[p + 1]
        _lib.check(lib.cmpy_transpose_pull_acc(_lib.ptr(y[r0 * nd:r1 * nd]), r1 - r0, nd, r0, nu, world,
                                               bounds, peers_y, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = orc.hubbard_matvec_free(up, dn, nb, kw["inter"], -kw["mu"], kw["hop"], xh, width=L)
    got = y.cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    full = model.hamilton_operator(nu_f, nd_f).matvec(x).cpu().numpy()
    assert np.abs(got - full).max() / np.abs(ref).max() < 1e-12


@pytest.mark.parametrize("L,nu_f,nd_f,nbfn", [
    (8, 4, 4, lambda: orc.chain_neighbors(8)),
    (10, 5, 5, lambda: orc.square_neighbors(2, 5)),
    (12, 6, 6, lambda: orc.square_neighbors(4, 3)),
    (18, 1, 9, lambda: orc.chain_neighbors(18)),     # rows of more than 16 sites: long-row kernel (the C5 path)
])
def test_c_dist_calls_world1(L, nu_f, nd_f, nbfn):
    """cmpy_dist_create / cmpy_hv_apply_sharded / cmpy_dist_allreduce_sum / cmpy_lanczos_sharded with a
    world of one rank (peer tables = local buffers): the C choreography, the library's own barrier and
    all-reduce kernels, the scaled accumulation of the row engine and of the pull kernel, and the
    device-scalar Lanczos driver -- against the oracle and the single-GPU fused Lanczos."""
    import torch
    from cmpy_b200 import _lib
    from cmpy_b200.exactdiag import lanczos_run
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.operators import SectorHamiltonOperator

    lib = _lib.lib()
    nb = nbfn()
    kw = dict(inter=4.0, mu=2.0, hop=1.0)
    model = HubbardModel(L, nb, **kw)
    spec = model._operator_spec()
    up, dn = orc.enumerate_states(L, nu_f), orc.enumerate_states(L, nd_f)
    nu, nd = len(up), len(dn)
    op_main = SectorHamiltonOperator(L, up, dn, spec["bonds"], spec["hops"], spec["eps"], spec["u"],
                                     spec["sign_width"])
    zeros = np.zeros(L)
    op_t = SectorHamiltonOperator(L, dn, up, spec["bonds"], spec["hops"], zeros, zeros, spec["sign_width"])
    xt = torch.zeros(nu * nd, dtype=torch.float64, device="cuda")
    yt = torch.zeros_like(xt)
    ctl = torch.zeros(int(lib.cmpy_dist_ctl_bytes()) // 8, dtype=torch.float64, device="cuda")
    one = lambda t: (ctypes.c_void_p * 1)(t.data_ptr())
    handle = ctypes.c_void_p()
    _lib.check(lib.cmpy_dist_create(op_main.handle, op_t.handle, 1, 0, one(xt), one(yt), one(ctl),
                                    ctypes.byref(handle)))
    try:
        xh = np.random.default_rng(9).standard_normal(nu * nd)
        x = torch.from_numpy(xh).cuda()
        y = torch.full_like(x, float("nan"))
        _lib.check(lib.cmpy_hv_apply_sharded(handle, _lib.ptr(x), _lib.ptr(y), 0, _lib.stream_ptr()))
        ref = orc.hubbard_matvec_free(up, dn, nb, kw["inter"], -kw["mu"], kw["hop"], xh, width=L)
        assert np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-12
        y2 = x.clone()
        _lib.check(lib.cmpy_hv_apply_sharded(handle, _lib.ptr(x), _lib.ptr(y2), 1, _lib.stream_ptr()))
        assert np.abs(y2.cpu().numpy() - (ref + xh)).max() / np.abs(ref).max() < 1e-12
        part = torch.tensor([1.25, -3.5], dtype=torch.float64, device="cuda")
        tot = torch.zeros(2, dtype=torch.float64, device="cuda")
        for _ in range(3):   # both slot parities
            _lib.check(lib.cmpy_dist_allreduce_sum(handle, _lib.ptr(part), _lib.ptr(tot), _lib.stream_ptr()))
            _lib.check(lib.cmpy_dist_barrier(handle, _lib.stream_ptr()))
        assert tot.cpu().tolist() == [1.25, -3.5]
        # sharded Lanczos driver == single-GPU fused Lanczos (same start vector), coefficient by coefficient
        h = model.hamilton_operator(nu_f, nd_f)
        res = lanczos_run(h, x, maxit=60, tol=0.0, check_every=60)
        r, w = x.clone(), torch.empty_like(x)
        alpha = np.zeros(80); beta = np.zeros(81)
        nit, e0 = ctypes.c_int(0), ctypes.c_double(0.0)
        rc = lib.cmpy_lanczos_sharded(handle, _lib.ptr(r), _lib.ptr(w), 60, 0.0, 60,
                                      alpha.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      beta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      ctypes.byref(nit), ctypes.byref(e0), _lib.stream_ptr())
        if rc == _lib.CMPY_ERR_UNSUPPORTED:   # sector outside the row engine (e.g. more than 4 LH bonds):
            assert b"scaled accumulation" in lib.cmpy_last_error()   # the Python recurrence is the documented fallback
            return
        assert rc in (_lib.CMPY_OK, _lib.CMPY_ERR_NOT_CONVERGED), lib.cmpy_last_error()
        m = min(nit.value, res.nit, 40)
        assert m >= 20
        assert np.allclose(alpha[:m], res.alpha[:m], rtol=0, atol=1e-9 * max(1.0, np.abs(res.alpha).max()))
        assert np.allclose(beta[:m + 1], res.beta[:m + 1], rtol=0, atol=1e-9 * max(1.0, np.abs(res.beta).max()))
        # and converges to the ground-state energy of the dense / sparse reference
        r = x.clone()
        rc = lib.cmpy_lanczos_sharded(handle, _lib.ptr(r), _lib.ptr(w), 70, 1e-11, 10,
                                      alpha.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      beta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      ctypes.byref(nit), ctypes.byref(e0), _lib.stream_ptr())
        ref_run = lanczos_run(h, x, maxit=400, tol=1e-12)
        if rc == _lib.CMPY_OK:
            assert abs(e0.value - ref_run.e0) < 1e-9
    finally:
        lib.cmpy_dist_destroy(handle)


def test_c_example_runs():
    """examples/e0_from_c.c on the GPU: E0 of the 8-site chain from a plain-C client of the ABI."""
    import subprocess
    import tempfile

    from cmpy_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "e0")
        subprocess.run(["gcc", "-O1", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(root, "examples", "e0_from_c.c"), "-L", os.path.dirname(_lib.LIB_PATH),
                        "-lcmpy_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm",
                        "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-o", exe], check=True)
        res = subprocess.run([exe], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "E0 = -20.2358069991" in res.stdout
