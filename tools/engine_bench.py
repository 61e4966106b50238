"""Engine 0 vs engine 2 of the class-major kernel on one GPU: parity between the variants and
CUDA-event timings of the full H.v and of the dn-only row-slab pass.
`python tools/engine_bench.py [reps]` -> JSON lines on stdout and gpurun_out/engine_bench.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cmpy_b200.models import HubbardModel  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = sys.argv[2] if len(sys.argv) > 2 else "bench"   # "ncu": one launch per variant, for a profiler run
out_path = os.path.join(ROOT, "gpurun_out", "engine_bench.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
out_f = open(out_path, "a")


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    out_f.write(line + "\n")
    out_f.flush()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def square(nx, ny):
    return [[nx * r + c, nx * r + c + 1] for r in range(ny) for c in range(nx - 1)] + \
           [[nx * r + c, nx * (r + 1) + c] for r in range(ny - 1) for c in range(nx)]


def run(name, L, nb, nu, nd, full_variants, slab_variants):
    h = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0).hamilton_operator(nu, nd)
    n = h.shape[0]
    num_up = len(h.up_states)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty_like(x)
    rec = {"workload": name, "dim": n, "bytes_algorithmic": 16 * n}
    ref = None
    for v in full_variants:
        try:
            h.set_variant(v)
            h.apply(x, out=y)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            rec[f"full_v{v}"] = f"unavailable: {exc}"
            continue
        if ref is None:
            ref = y.clone()
        else:
            rec[f"full_v{v}_relerr"] = float((y - ref).abs().max() / ref.abs().max())
        ms = timed(lambda: h.apply(x, out=y), reps)
        rec[f"full_v{v}_ms"] = ms
        rec[f"full_v{v}_gbs"] = 16 * n / ms / 1e6
    ref = None
    for v in slab_variants:
        try:
            h.set_variant(v)
            h.apply_rows(x, 0, num_up, out=y)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            rec[f"dn_v{v}"] = f"unavailable: {exc}"
            continue
        if ref is None:
            ref = y.clone()
        else:
            rec[f"dn_v{v}_relerr"] = float((y - ref).abs().max() / ref.abs().max())
        ms = timed(lambda: h.apply_rows(x, 0, num_up, out=y), reps)
        rec[f"dn_v{v}_ms"] = ms
        rec[f"dn_v{v}_gbs"] = 16 * n / ms / 1e6
    h.set_variant(0)
    emit(rec)
    del h, x, y
    torch.cuda.empty_cache()


def ncu_launches():
    """One dn-only launch of engine 0 (variant 5) and engine 2 (variant 9) on config C4."""
    h = HubbardModel(16, square(4, 4), inter=4.0, mu=2.0, hop=1.0).hamilton_operator(8, 8)
    x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    for v in (5, 9):
        h.set_variant(v)
        h.apply_rows(x, 0, len(h.up_states), out=y)
    torch.cuda.synchronize()


if __name__ == "__main__":
    torch.cuda.set_device(0)
    if mode == "ncu":
        ncu_launches()
        sys.exit(0)
    if mode == "threads":  # engine 2 with 512 / 768 / 1024-thread CTAs and 16..64 work pieces
        for nt, npc in ((512, 16), (512, 32), (768, 24), (896, 28), (896, 56), (1024, 32)):
            os.environ["CMPY_CLS2_THREADS"] = str(nt)
            os.environ["CMPY_CLS_PIECES"] = str(npc)
            run(f"c4_square4x4_t{nt}_p{npc}", 16, square(4, 4), 8, 8, [9], [5, 9])
        os.environ["CMPY_CLS2_THREADS"] = "512"
        os.environ["CMPY_CLS_PIECES"] = "32"
        run("chain16_t512_p32", 16, [[i, i + 1] for i in range(15)], 8, 8, [0, 9], [5, 9])
        sys.exit(0)
    if mode == "pieces":   # engine 2 with different numbers of work pieces per phase
        for npc in (32, 48, 64, 96, 128):
            os.environ["CMPY_CLS_PIECES"] = str(npc)
            run(f"c4_square4x4_pieces{npc}", 16, square(4, 4), 8, 8, [9], [5, 9])
        sys.exit(0)
    # BASELINE config C4 and the 16-site chain: 0 = default (segment kernel), 5 = class-major engine 0,
    # 9 = class-major engine 2
    run("c4_square4x4", 16, square(4, 4), 8, 8, [0, 5, 9], [5, 9])
    run("chain16", 16, [[i, i + 1] for i in range(15)], 8, 8, [0, 5, 9], [5, 9])
    # long rows (20-site chain, 45 up rows x 184756 dn strings): 1 = global gather, 8 / 10 = long-row engines 0 / 2
    run("chain20_slab", 20, [[i, i + 1] for i in range(19)], 2, 10, [], [1, 8, 10])
    out_f.close()
