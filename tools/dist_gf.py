"""Sharded spectral function on N GPUs (run under torchrun): G_{0,DN}(z) of the Hubbard model by
sharded Lanczos + continued fraction, checked against the single-GPU path and timed.
usage: dist_gf.py [c2|c4|chain14]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.basis import DN
from cmpy_b200.dist import gf_continued_fraction_sharded
from cmpy_b200.exactdiag import gf_continued_fraction

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
L, nb = {"c2": (12, orc.chain_neighbors(12)), "chain14": (14, orc.chain_neighbors(14)),
         "c4": (16, orc.square_neighbors(4, 4))}[wl]
model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
z = np.linspace(-6, 6, 1001) + 0.05j
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 600
gf_continued_fraction_sharded(HubbardModel(8, orc.chain_neighbors(8), inter=4.0, mu=2.0, hop=1.0), z[:8], num_coeffs=20)  # warm-up
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
g, info = gf_continued_fraction_sharded(model, z, pos=0, num_coeffs=NC, return_info=True)
torch.cuda.synchronize(); dist.barrier(); t_sh = time.time() - t0
out = dict(workload=wl, n_gpus=world, seconds=t_sh, e0=info["e0"], norms=info["norms"], nit=info["nit"],
           gs_iterations=info["gs_iterations"], g_mid=[float(g[500].real), float(g[500].imag)],
           sumrule=float(-np.trapezoid(g.imag, z.real) / np.pi))
if rank == 0 and (wl != "c4" or world <= 2):
    t0 = time.time()
    gref = gf_continued_fraction(model, z, pos=0, sigma=DN, num_coeffs=NC)
    torch.cuda.synchronize()
    out["single_gpu_seconds"] = time.time() - t0
    out["max_abs_diff_vs_single_gpu"] = float(np.abs(g - gref).max())
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
