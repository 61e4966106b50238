"""GPU checks of the opt-in code paths that have not been measured yet (run with
CMPY_EXPERIMENTAL=1): the partial pull transpose of the chunked second half of the sharded H.v,
emulated on one GPU with several 'virtual ranks' whose slabs live in local buffers."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CMPY_EXPERIMENTAL") != "1", reason="set CMPY_EXPERIMENTAL=1")]


@pytest.mark.parametrize("nparts", [1, 2, 3, 5])
def test_partial_pull_with_virtual_ranks(nparts):
    import torch
    from cmpy_b200 import _lib

    L = _lib.lib()
    world, nd, nu_total, row0, nrows = 3, 131, 70, 5, 50
    cb = [0, 40, 90, 131]
    g = torch.Generator(device="cuda").manual_seed(4)
    yts = [torch.randn((cb[q + 1] - cb[q]) * nu_total, dtype=torch.float64, device="cuda", generator=g)
           for q in range(world)]
    y0 = torch.randn(nrows * nd, dtype=torch.float64, device="cuda", generator=g)
    ref = y0.clone().view(nrows, nd)
    for q in range(world):
        blk = yts[q].view(cb[q + 1] - cb[q], nu_total)[:, row0:row0 + nrows]     # [c - cb[q], r]
        ref[:, cb[q]:cb[q + 1]] += blk.t()
    peers = (ctypes.c_void_p * world)(*[t.data_ptr() for t in yts])
    bounds = (ctypes.c_int64 * (world + 1))(*cb)
    y = y0.clone()
    for part in range(nparts):
        _lib.check(L.cmpy_transpose_pull_acc_part(_lib.ptr(y), nrows, nd, row0, nu_total, world, bounds, peers,
                                                  part, nparts, 7 if part % 2 else 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(y.view(nrows, nd), ref)
    # one part alone touches exactly its columns
    y = y0.clone()
    _lib.check(L.cmpy_transpose_pull_acc_part(_lib.ptr(y), nrows, nd, row0, nu_total, world, bounds, peers,
                                              0, nparts, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    touched = (y.view(nrows, nd) != y0.view(nrows, nd)).any(dim=0).cpu().numpy()
    expect = np.zeros(nd, dtype=bool)
    for q in range(world):
        n = cb[q + 1] - cb[q]
        expect[cb[q] + n * 0 // nparts: cb[q] + n * 1 // nparts] = True
    assert (touched == expect).all()
