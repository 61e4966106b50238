"""Developer micro-benchmark of the generation-3 row engine (variant 11) against the older kernels:
dn-only slab pass and full H.v, CUDA events, parity against the seg kernel printed beside the time."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.exactdiag import lanczos_run


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(name, L, nb, nu, nd, n=10):
    h = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0).hamilton_operator(nu, nd)
    dim = h.shape[0]
    x = torch.randn(dim, dtype=torch.float64, device="cuda"); x /= x.norm()
    y = torch.empty_like(x); ref = torch.empty_like(x)
    out = {"config": name, "dim": dim}
    h.set_variant(4); h.apply_rows(x, 0, len(h.up_states), out=ref)
    for v in (11, 5, 4):
        try:
            h.set_variant(v)
            ms = timeit(lambda: h.apply_rows(x, 0, len(h.up_states), out=y), n=n)
            err = float((y - ref).abs().max() / ref.abs().max())
            out[f"dn_only_v{v}_ms"] = ms
            print(f"{name} dn-only variant={v}: {ms:.4f} ms  {16*dim/ms/1e6:.1f} GB/s alg  relerr vs v4 {err:.2e}", flush=True)
        except Exception as e:
            print(name, "dn-only variant", v, "failed:", e, flush=True)
    h.set_variant(4); h.apply(x, out=ref)
    for v in (11, 4, 5):
        try:
            h.set_variant(v)
            ms = timeit(lambda: h.apply(x, out=y), n=n)
            err = float((y - ref).abs().max() / ref.abs().max())
            out[f"full_v{v}_ms"] = ms
            print(f"{name} full variant={v}: {ms:.4f} ms  {16*dim/ms/1e6:.1f} GB/s alg  relerr vs v4 {err:.2e}", flush=True)
        except Exception as e:
            print(name, "full variant", v, "failed:", e, flush=True)
    for v in (0, 11):
        h.set_variant(v)
        lanczos_run(h, None, maxit=10, tol=1e-10)
        torch.cuda.synchronize(); t0 = time.time()
        res = lanczos_run(h, None, maxit=400, tol=1e-10)
        torch.cuda.synchronize()
        out[f"lanczos_v{v}_s"] = time.time() - t0
        print(f"{name} lanczos variant={v}: e0={res.e0:.12f} it={res.nit} conv={res.converged} {time.time()-t0:.3f}s", flush=True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4", "c16"]
    if "c2" in which: run("C2 L=12", 12, orc.chain_neighbors(12), 6, 6, n=50)
    if "c14" in which: run("L=14", 14, orc.chain_neighbors(14), 7, 7, n=20)
    if "c4" in which: run("C4 4x4", 16, orc.square_neighbors(4, 4), 8, 8, n=10)
    if "c16" in which: run("chain16", 16, orc.chain_neighbors(16), 8, 8, n=10)
