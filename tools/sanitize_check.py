"""Small-sector walk through every shipped kernel family, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck) on one GPU:

    compute-sanitizer --tool memcheck  python tools/sanitize_check.py
    compute-sanitizer --tool racecheck python tools/sanitize_check.py
    compute-sanitizer --tool synccheck python tools/sanitize_check.py

Sectors are small (at most 924 x 924) so that the instrumented run finishes in a minute or two; results are
still compared with the numpy oracle.  Test infrastructure only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle_np as orc
from cmpy_b200.models import HubbardModel, HeisenbergModel
from cmpy_b200.exactdiag import lanczos_run, gf_continued_fraction
from cmpy_b200.dist import ShardedHubbardOperator, lanczos_sharded


def relerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def hubbard_case(L, nb, nu, nd, variants):
    kw = dict(inter=4.0, mu=2.0, hop=1.0)
    model = HubbardModel(L, nb, **kw)
    h = model.hamilton_operator(nu, nd)
    up, dn = orc.enumerate_states(L, nu), orc.enumerate_states(L, nd)
    x = np.random.default_rng(3).standard_normal(h.shape[0])
    ref = orc.hubbard_matvec_free(up, dn, nb, 4.0, -2.0, 1.0, x, width=L)
    ran = []
    for v in variants:
        try:
            h.set_variant(v)
            y = h.matvec(x)
        except RuntimeError as exc:
            assert "variant" in str(exc), exc
            continue
        assert relerr(y, ref) < 1e-12, (L, nu, nd, v)
        ran.append(v)
    h.set_variant(0)
    xd = torch.from_numpy(x).cuda()
    # row-slab entry point (default: row engine), plain and accumulating
    out = torch.zeros_like(xd)
    h.apply_rows(xd, 0, len(up), out=out)
    h.apply_rows(xd, 0, len(up), out=out, accumulate=True)
    res = lanczos_run(h, None, maxit=400, tol=1e-11)         # fused Lanczos kernels + CUDA graph replay
    # sharded operator in a world of one rank: C choreography, peer transposes, device barrier / all-reduce
    sh = ShardedHubbardOperator(model, nu, nd)
    got = sh.apply_local(xd)
    assert relerr(got.cpu().numpy(), ref) < 1e-12
    e0 = lanczos_sharded(sh, maxit=400, tol=1e-11)[0]
    assert abs(e0 - res.e0) < 1e-8, (e0, res.e0)
    print(f"hubbard L={L} ({nu},{nd}): variants {ran} ok, lanczos e0 {res.e0:.10f} / sharded {e0:.10f}", flush=True)


def main():
    chain = lambda L, per=False: orc.chain_neighbors(L, per)
    hubbard_case(8, chain(8), 4, 4, (1, 2, 3, 4, 5, 6, 7, 11))
    hubbard_case(10, orc.square_neighbors(2, 5), 5, 5, (1, 2, 3, 4, 5, 11))
    hubbard_case(12, orc.square_neighbors(4, 3), 6, 6, (1, 3, 4, 5, 11))
    hubbard_case(10, chain(10, True), 4, 6, (1, 3, 11))
    # long rows (more than 16 sites per string): sub-row launches of the class-major kernel
    m = HubbardModel(18, chain(18), inter=4.0, mu=2.0, hop=1.0)
    h = m.hamilton_operator(1, 9)
    x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda")
    h.set_variant(1); y1 = h.apply(x).clone()
    h.set_variant(0); y0 = h.apply(x)
    out = torch.empty_like(x); h.apply_rows(x, 0, 18, out=out)
    assert float((y1 - y0).abs().max()) < 1e-12 * float(y1.abs().max())
    print("hubbard L=18 (1,9) long rows ok", flush=True)
    # Heisenberg / XXZ kernels and the Lanczos + continued-fraction Green's function
    from refshim import ChainStandIn
    for N in (10, 12):
        hm = HeisenbergModel(ChainStandIn(N, periodic=True), j=0.9, jz=1.1).hamilton_operator(s=0)
        xs = np.random.default_rng(5).standard_normal(hm.shape[0])
        ys = hm.matvec(xs)
        assert np.isfinite(ys).all()
        lanczos_run(hm, None, maxit=30, tol=1e-10)
    print("heisenberg ok", flush=True)
    model = HubbardModel(8, chain(8), inter=4.0, mu=2.0, hop=1.0)
    z = np.linspace(-6, 6, 101) + 0.05j
    g = gf_continued_fraction(model, z, pos=0, n_up=4, n_dn=4, num_coeffs=80)
    assert np.isfinite(np.asarray(g)).all()
    print("gf ok", flush=True)
    torch.cuda.synchronize()
    print("sanitize_check done", flush=True)


if __name__ == "__main__":
    main()
