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


@pytest.mark.parametrize("world,nu,nd", [(2, 70, 131), (3, 70, 131), (5, 33, 64), (8, 257, 40)])
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
