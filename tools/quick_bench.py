"""Developer micro-benchmark (not the contract bench): times H.v variants with CUDA events."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.exactdiag import lanczos_run
from cmpy_b200 import _lib

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def run(name, L, nb, nu, nd, variants=(3, 4, 5, 6, 7), n=10):
    t0 = time.time()
    h = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0).hamilton_operator(nu, nd)
    tb = time.time() - t0
    dim = h.shape[0]
    x = torch.randn(dim, dtype=torch.float64, device="cuda"); x /= x.norm()
    y = torch.empty_like(x)
    for v in variants:
        try:
            h.set_variant(v)
            ms = timeit(lambda: h.apply(x, out=y), n=n)
            print(f"{name} dim={dim} variant={v}: {ms:.4f} ms  {16*dim/ms/1e6:.1f} GB/s algorithmic  (build {tb:.2f}s)", flush=True)
        except Exception as e:
            print(name, "variant", v, "failed:", e, flush=True)
    for v in (4, 5, 6, 7):   # dn-only pass (row slab entry point, no up hops)
        try:
            h.set_variant(v)
            ms = timeit(lambda: h.apply_rows(x, 0, len(h.up_states), out=y), n=n)
            print(f"{name} dn-only variant={v}: {ms:.4f} ms  {16*dim/ms/1e6:.1f} GB/s algorithmic", flush=True)
        except Exception as e:
            print(name, "dn-only variant", v, "failed:", e, flush=True)
    h.set_variant(0)
    t0 = time.time()
    res = lanczos_run(h, None, maxit=400, tol=1e-10)
    torch.cuda.synchronize()
    print(f"{name} lanczos: e0={res.e0:.12f} it={res.nit} conv={res.converged} {time.time()-t0:.3f}s", flush=True)
    return h

if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4"]
    # copy bandwidth reference
    a = torch.empty(2**28, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a))
    print(f"torch copy 2 GiB: {ms:.3f} ms -> {2*a.numel()*8/ms/1e6:.0f} GB/s", flush=True)
    del a, b
    if "c1" in which: run("C1 L=8", 8, orc.chain_neighbors(8), 4, 4, n=50)
    if "c2" in which: run("C2 L=12", 12, orc.chain_neighbors(12), 6, 6, n=50)
    if "c14" in which: run("L=14", 14, orc.chain_neighbors(14), 7, 7, n=20)
    if "c4" in which: run("C4 4x4", 16, orc.square_neighbors(4, 4), 8, 8, n=10)
    if "c16" in which: run("chain16", 16, orc.chain_neighbors(16), 8, 8, n=10)
