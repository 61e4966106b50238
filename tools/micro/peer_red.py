"""2-GPU microbenchmark (torchrun): remote stores vs remote red.add.f64 into the peer's symmetric buffer.
   torchrun --nproc-per-node 2 tools/micro/peer_red.py"""
import ctypes
import os
import subprocess

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

here = os.path.dirname(os.path.abspath(__file__))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
so = os.path.join(here, "libpeer_red.so")
if rank == 0 and not os.path.exists(so):
    subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                    "-o", so, os.path.join(here, "peer_red.cu")], check=True)
dist.barrier()
L = ctypes.CDLL(so)
L.run_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
L.run_tile.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
rows, ld = 6400, 12800                     # 655 MB
n = rows * ld
buf = symm.empty(n, dtype=torch.float64, device="cuda")
h = symm.rendezvous(buf, dist.group.WORLD)
buf.zero_()
src = torch.randn(n, dtype=torch.float64, device="cuda")
peer = int(h.buffer_ptrs[(rank + 1) % world])
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for name, fn in [
    ("store  linear, 148x8 CTAs", lambda: L.run_copy(peer, src.data_ptr(), n, 0, 148 * 8, st)),
    ("store  linear,  32x8 CTAs", lambda: L.run_copy(peer, src.data_ptr(), n, 0, 32 * 8, st)),
    ("red.sys linear, 148x8 CTAs", lambda: L.run_copy(peer, src.data_ptr(), n, 1, 148 * 8, st)),
    ("red.sys linear,  32x8 CTAs", lambda: L.run_copy(peer, src.data_ptr(), n, 1, 32 * 8, st)),
    ("red.gpu linear,  32x8 CTAs", lambda: L.run_copy(peer, src.data_ptr(), n, 2, 32 * 8, st)),
    ("store  tiles 32x128, 32x8 CTAs", lambda: L.run_tile(peer, src.data_ptr(), rows, ld, 0, 32 * 8, st)),
    ("red.sys tiles 32x128, 32x8 CTAs", lambda: L.run_tile(peer, src.data_ptr(), rows, ld, 1, 32 * 8, st)),
    ("red.sys tiles 32x128, 16x8 CTAs", lambda: L.run_tile(peer, src.data_ptr(), rows, ld, 1, 16 * 8, st)),
    ("red.sys LOCAL linear, 148x8 CTAs", lambda: L.run_copy(buf.data_ptr(), src.data_ptr(), n, 1, 148 * 8, st)),
]:
    ms = timed(fn)
    if rank == 0:
        print(f"{name:36s} {ms:8.3f} ms  {n * 8 / ms / 1e6:8.1f} GB/s per direction", flush=True)
# correctness of the remote reduction: buf was zeroed, then received stores and k reds of the peer's src
dist.barrier()
dist.destroy_process_group()
