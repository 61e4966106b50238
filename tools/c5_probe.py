import os, sys, time, traceback
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
from cmpy_b200.models import HubbardModel
from cmpy_b200.dist import ShardedHubbardOperator, ShardPlan
L = 20
nb = [[i, i + 1] for i in range(L - 1)]
try:
    t0 = time.time()
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    sector = model.basis.get_sector(10, 10)
    print("sector", len(sector.up_states), time.time() - t0, flush=True)
    from cmpy_b200.operators import SectorHamiltonOperator
    spec = model._operator_spec()
    up, dn = np.asarray(sector.up_states), np.asarray(sector.dn_states)
    op = SectorHamiltonOperator(L, up, dn, spec["bonds"], spec["hops"], spec["eps"], spec["u"], spec["sign_width"])
    print("operator built", time.time() - t0, flush=True)
    nrows = 256
    x = torch.randn(nrows * len(dn), dtype=torch.float64, device="cuda")
    y = op.apply_rows(x, 1000, nrows)
    torch.cuda.synchronize()
    print("apply_rows ok", float(y.abs().max()), flush=True)
    import torch.distributed._symmetric_memory as symm
    n = 23095 * 184756
    t = symm.empty(n, dtype=torch.float64, device="cuda")
    h = symm.rendezvous(t, dist.group.WORLD)
    print("symm 34GB ok", h.buffer_ptrs, flush=True)
    t2 = symm.empty(n, dtype=torch.float64, device="cuda")
    h2 = symm.rendezvous(t2, dist.group.WORLD)
    print("symm 2x34GB ok", flush=True)
    v = torch.randn(n, dtype=torch.float64, device="cuda"); w = torch.zeros_like(v)
    print("norm", float(v[:1 << 30].norm()), flush=True)
    # one slab-sized apply_rows (flat kernel at full slab scale)
    torch.cuda.synchronize(); t1 = time.time()
    op.apply_rows(v, 0, 23095, out=w)
    torch.cuda.synchronize(); print("slab apply_rows s", time.time() - t1, flush=True)
    t1 = time.time(); op.apply_rows(v, 0, 23095, out=w); torch.cuda.synchronize(); print("slab apply_rows s (2nd)", time.time() - t1, flush=True)
    w.mul_(-0.5); w.addcmul_(v, torch.tensor(0.3, device="cuda", dtype=torch.float64), value=-1.0); w.div_(torch.tensor(2.0, device="cuda", dtype=torch.float64))
    torch.cuda.synchronize(); print("vector ops ok", flush=True)
except Exception:
    traceback.print_exc()
dist.destroy_process_group()
