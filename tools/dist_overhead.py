"""Where does a sharded Lanczos iteration spend its time? (torchrun, diagnostic)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.dist import ShardedHubbardOperator
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
model = HubbardModel(L, orc.chain_neighbors(L), inter=4.0, mu=2.0, hop=1.0)
op = ShardedHubbardOperator(model, L // 2, L // 2)
v = torch.randn(op.local_size, dtype=torch.float64, device="cuda"); w = torch.zeros_like(v)
a = torch.zeros((), dtype=torch.float64, device="cuda")
def timeit(name, fn, n=200):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e6
    if rank == 0: print(f"{name:28s} {dt:9.1f} us", flush=True)
timeit("apply_local", lambda: op.apply_local(v, out=w, accumulate=True))
timeit("symm barrier", lambda: op._h_xt.barrier(channel=0))
timeit("torch.dot", lambda: torch.dot(v, w))
timeit("all_reduce 0-dim", lambda: dist.all_reduce(a))
timeit("dot + all_reduce", lambda: dist.all_reduce(torch.dot(v, w)))
timeit("mul_", lambda: w.mul_(a))
timeit("addcmul_", lambda: w.addcmul_(v, a, value=-1.0))
timeit("div_", lambda: w.div_(torch.ones((), dtype=torch.float64, device="cuda")))
alphas = torch.zeros(100, dtype=torch.float64, device="cuda")
timeit("alphas[j] = a", lambda: alphas.__setitem__(3, a))
timeit("push only", lambda: op.backend.apply_rows(v, op.plan.rows()[0], op.plan.nrows, w))
dist.destroy_process_group()
