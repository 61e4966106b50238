"""ncu target: a few launches of one H.v variant on one config (dn-only slab pass or full H.v)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import oracle_np as orc
from cmpy_b200.models import HubbardModel

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 11
mode = sys.argv[3] if len(sys.argv) > 3 else "dn"
L, nb, nu, nd = {"c4": (16, orc.square_neighbors(4, 4), 8, 8), "c16": (16, orc.chain_neighbors(16), 8, 8),
                 "c2": (12, orc.chain_neighbors(12), 6, 6)}[cfg]
h = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0).hamilton_operator(nu, nd)
x = torch.randn(h.shape[0], dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
h.set_variant(variant)
for _ in range(3):
    if mode == "dn":
        h.apply_rows(x, 0, len(h.up_states), out=y)
    else:
        h.apply(x, out=y)
torch.cuda.synchronize()
print("done")
