"""Timings of every BASELINE.json config that fits one GPU (C1-C4): H.v, Lanczos E0 to 1e-10,
continued-fraction G(omega) on the reference grid.  Prints one JSON object per config."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle_np as orc
from refshim import ChainStandIn
from cmpy_b200.models import HubbardModel, HeisenbergModel
from cmpy_b200.exactdiag import lanczos_run, gf_continued_fraction
from cmpy_b200 import _lib

PEAK = 6541.1
which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]

def time_hv(h, n=20):
    x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda"); x /= x.norm()
    y = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if 8 * h.shape[0] < 200e6 else None
    for _ in range(3):
        h.apply(x, out=y)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); h.apply(x, out=y); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n

def hubbard(name, L, nb, n, gf=True):
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    t0 = time.time(); h = model.hamilton_operator(n, n); torch.cuda.synchronize(); tb = time.time() - t0
    dim = h.shape[0]
    ms = time_hv(h)
    t0 = time.time(); res = lanczos_run(h, None, maxit=1000, tol=1e-10, check_every=10, want_vector=gf)
    torch.cuda.synchronize(); tl = time.time() - t0
    out = dict(config=name, dim=dim, build_s=tb, hv_ms=ms, hv_gbs_algorithmic=16 * dim / ms / 1e6,
               hv_frac_of_measured_hbm=16 * dim / ms / 1e6 / PEAK, lanczos_s=tl, lanczos_it=res.nit, e0=res.e0)
    if gf:
        z = np.linspace(-6, 6, 1001) + 0.05j
        t0 = time.time()
        g, info = gf_continued_fraction(model, z, pos=0, gs=(res.e0, res.vector), num_coeffs=600, return_info=True)
        torch.cuda.synchronize()
        out.update(gf_s=time.time() - t0, gf_nit=info["nit"], gf_norms=info["norms"],
                   gf_sumrule=float(-np.trapz(np.asarray(g).imag, z.real) / np.pi), g_mid=[float(np.asarray(g)[500].real), float(np.asarray(g)[500].imag)])
    print(json.dumps(out), flush=True)

if "c1" in which: hubbard("C1 hubbard chain L=8 (4,4)", 8, orc.chain_neighbors(8), 4)
if "c2" in which: hubbard("C2 hubbard chain L=12 (6,6)", 12, orc.chain_neighbors(12), 6)
if "c3" in which:
    N = 32
    model = HeisenbergModel(ChainStandIn(N), j=1.0, jz=1.0)
    t0 = time.time(); h = model.hamilton_operator(s=0); torch.cuda.synchronize(); tb = time.time() - t0
    dim = h.shape[0]
    ms = time_hv(h, n=10)
    t0 = time.time(); res = lanczos_run(h, None, maxit=1000, tol=1e-10, check_every=10)
    torch.cuda.synchronize(); tl = time.time() - t0
    print(json.dumps(dict(config="C3 heisenberg chain N=32 Sz=0", dim=dim, build_s=tb, hv_ms=ms,
                          hv_gbs_algorithmic=16 * dim / ms / 1e6, hv_frac_of_measured_hbm=16 * dim / ms / 1e6 / PEAK,
                          lanczos_s=tl, lanczos_it=res.nit, e0=res.e0, converged=bool(res.converged))), flush=True)
    del h
    torch.cuda.empty_cache()
if "c4" in which: hubbard("C4 hubbard 4x4 (8,8)", 16, orc.square_neighbors(4, 4), 8, gf=True)
if "c16" in which: hubbard("hubbard chain L=16 (8,8)", 16, orc.chain_neighbors(16), 8, gf=False)
