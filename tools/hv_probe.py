"""Developer probe: a few H.v launches of one variant (run under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import oracle_np as orc
from cmpy_b200.models import HubbardModel
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = {"c4": (16, orc.square_neighbors(4, 4), 8, 8), "c2": (12, orc.chain_neighbors(12), 6, 6),
       "c16": (16, orc.chain_neighbors(16), 8, 8)}[which]
h = HubbardModel(cfg[0], cfg[1], inter=4.0, mu=2.0, hop=1.0).hamilton_operator(cfg[2], cfg[3])
h.set_variant(variant)
x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
dn_only = len(sys.argv) > 4 and sys.argv[4] == "dn"
for _ in range(n):
    if dn_only:
        h.apply_rows(x, 0, len(h.up_states), out=y)
    else:
        h.apply(x, out=y)
torch.cuda.synchronize()
print("done", which, variant, float(y[0]))
