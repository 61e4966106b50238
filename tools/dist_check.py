"""Multi-GPU parity check (run under torchrun): sharded H.v slab == slab of the single-GPU H.v."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
import oracle_np as orc
from cmpy_b200.models import HubbardModel
from cmpy_b200.dist import ShardedHubbardOperator

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cases = [(12, orc.chain_neighbors(12), 6, 6), (12, orc.square_neighbors(4, 3), 5, 7),
         (14, orc.chain_neighbors(14, True), 7, 7), (16, orc.square_neighbors(4, 4), 8, 8)]
if len(sys.argv) > 1 and sys.argv[1] == "c4":   # only the benchmarked sector (plus one small case)
    cases = [cases[0], cases[3]]
for L, nb, nu, nd in cases:
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    full = model.hamilton_operator(nu, nd)
    n = full.shape[0]
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y = full.matvec(x)
    for exchange in ("peer", "a2a"):
        sh = ShardedHubbardOperator(model, nu, nd, exchange=exchange)
        r0, r1 = sh.plan.rows()
        ndn = sh.plan.num_dn
        xl = x[r0 * ndn:r1 * ndn].clone()
        yl = sh.apply_local(xl)
        err = float((yl - y[r0 * ndn:r1 * ndn]).abs().max() / y.abs().max())
        t = torch.tensor([err], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # timing
        for _ in range(3):
            sh.apply_local(xl, out=yl)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sh.apply_local(xl, out=yl)
        e1.record(); dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"L={L} ({nu},{nd}) dim={n} world={world} exchange={exchange}: max rel err {float(t):.2e}, "
                  f"{float(ms):.3f} ms per H.v", flush=True)
        assert float(t) < 1e-12
        if exchange == "peer" and L >= 14:
            ph = sh.profile_phases(xl, yl)
            if rank == 0:
                gb = ph["nvlink_bytes_per_transpose"] / 1e9
                print("   phases (ms): " + ", ".join(f"{k} {v:.3f}" for k, v in ph.items() if k != "nvlink_bytes_per_transpose")
                      + f" | NVLink out {gb:.3f} GB/transpose -> push {gb / ph['push'] * 1e3:.0f} GB/s, pull {gb / ph['pull'] * 1e3:.0f} GB/s", flush=True)
        del sh, yl
    del full, x, y
    torch.cuda.empty_cache()
dist.barrier(); dist.destroy_process_group()
if rank == 0:
    print("dist_check ok")
