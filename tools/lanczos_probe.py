"""Developer probe: a few Lanczos iterations on a workload (run under ncu for a launch list)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.exactdiag import lanczos_run
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 10
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = {"c4": (16, orc.square_neighbors(4, 4), 8, 8), "c2": (12, orc.chain_neighbors(12), 6, 6)}[which]
h = HubbardModel(cfg[0], cfg[1], inter=4.0, mu=2.0, hop=1.0).hamilton_operator(cfg[2], cfg[3])
torch.cuda.synchronize(); t0 = time.time()
res = lanczos_run(h, None, maxit=nit, tol=0.0, resid_tol=0.0, check_every=nit, use_graph=bool(graph))
torch.cuda.synchronize()
print(which, "nit", res.nit, "time", time.time() - t0, "e0", res.e0)
