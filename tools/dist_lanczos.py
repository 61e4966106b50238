"""Sharded Lanczos E0 on N GPUs (run under torchrun): `dist_lanczos.py <workload> [U] [maxit]`.
workloads: c4 (4x4), chain16, chain18, chain20 (BASELINE config C5), half filling."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from cmpy_b200.models import HubbardModel
from cmpy_b200.dist import ShardedHubbardOperator, lanczos_sharded

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
U = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 400
if wl == "c4":
    L = 16
    nb = [[4 * r + c, 4 * r + c + 1] for r in range(4) for c in range(3)] + [[4 * r + c, 4 * r + c + 4] for r in range(3) for c in range(4)]
else:
    L = int(wl.replace("chain", ""))
    nb = [[i, i + 1] for i in range(L - 1)]
n = L // 2
t0 = time.time()
model = HubbardModel(L, nb, inter=U, mu=U / 2, hop=1.0)
op = ShardedHubbardOperator(model, n, n)
torch.cuda.synchronize(); dist.barrier()
t_build = time.time() - t0
dim = op.shape[0]
# H.v timing
x = torch.randn(op.local_size, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for _ in range(2):
    op.apply_local(x, out=y)
dist.barrier(); torch.cuda.synchronize()
e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0_.record()
for _ in range(reps):
    op.apply_local(x, out=y)
e1_.record(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0_.elapsed_time(e1_) / reps], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
hv_ms = float(ms)
phases = None
if op.exchange == "peer":
    ph = op.profile_phases(x, y, reps=3)
    t = torch.tensor([ph[k] for k in ("dn_pass", "push", "up_pass", "pull", "barrier")], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    phases = dict(zip(("dn_pass", "push", "up_pass", "pull", "barrier"), [float(v) for v in t]))
    phases["nvlink_bytes_per_transpose"] = ph["nvlink_bytes_per_transpose"]
del x, y
torch.cuda.empty_cache()
def cb(nit, e):
    if rank == 0:
        print(f"  it {nit:4d}  E0 = {e:.12f}  ({time.time() - t1:.1f} s)", flush=True)
dist.barrier(); torch.cuda.synchronize(); t1 = time.time()
verbose = os.environ.get("DIST_LANCZOS_VERBOSE", "0") == "1"   # the callback forces the Python recurrence
e0, al, be, nit, conv = lanczos_sharded(op, maxit=maxit, tol=1e-10, check_every=10, callback=cb if verbose else None)
torch.cuda.synchronize(); dist.barrier()
t_lz = time.time() - t1
mem = torch.cuda.max_memory_allocated() / 1e9
if rank == 0:
    p = op.plan
    out = dict(workload=wl, L=L, U=U, dim=dim, n_gpus=world, exchange=op.exchange, build_s=t_build,
               hv_ms=hv_ms, hv_algorithmic_gbs_per_gpu=16.0 * dim / world / (hv_ms * 1e-3) / 1e9,
               nvlink_out_gbs_per_gpu=p.bytes_out_per_hv() / (hv_ms * 1e-3) / 1e9,
               lanczos_path=getattr(op, "last_lanczos_path", "python"),
               lanczos_s=t_lz, iterations=nit, converged=bool(conv), e0=e0,
               ms_per_iteration=1e3 * t_lz / max(nit, 1), max_mem_gb=mem, phases_ms=phases)
    if U == 0.0:  # free fermions: E0 = 2 * sum of the n lowest levels of the hopping matrix (mu = 0)
        h = np.zeros((L, L))
        for i, j in nb:
            h[i, j] = h[j, i] = 1.0
        ev = np.linalg.eigvalsh(h)
        out["e0_exact_u0"] = 2.0 * float(ev[:n].sum())
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
