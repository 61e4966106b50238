#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Contract benchmark: `python bench.py --gpus N --steps K --warmup W [--impl reference]`.

Metric (BASELINE.json): H.v mat-vecs/s (+ achieved HBM GB/s) on the Hubbard 16-site
(4x4 open square lattice) half-filling sector, dim 165 636 900 (config C4); the Lanczos
E0 time to 1e-10 is reported beside it (`lanczos_e0`).

A "step" is one H.v over the whole sector.  N=1: one B200 holds the vector; N>1: the vector
is sharded by up-string across the N ranks (strong scaling, NCCL all-to-all transposes).
`value` is whole-job mat-vecs/s with the vector resident in HBM; `e2e` is the same metric
through the public operator call with pinned HOST buffers (H2D of x and D2H of y inside the
timed region).  `--impl reference` times the CPU port of the reference's H.v
(oracle/hv_oracle.c, OpenMP over all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hubbard_hv_matvecs_per_s"
UNIT = "matvec/s"
WORKLOADS = {
    # name: (num_sites, lattice, n_up, n_dn)
    "c4": (16, "square4x4", 8, 8),
    "c2": (12, "chain", 6, 6),
    "c1": (8, "chain", 4, 4),
    "chain14": (14, "chain", 7, 7),
}
PARAMS = dict(inter=4.0, mu=2.0, hop=1.0)  # SURVEY.md section 8(d) synthetic inputs


def neighbors_of(lattice, num_sites):
    if lattice == "chain":
        return [[i, i + 1] for i in range(num_sites - 1)]
    if lattice == "square4x4":
        nb = []
        for r in range(4):
            for c in range(4):
                i = 4 * r + c
                if c + 1 < 4:
                    nb.append([i, i + 1])
                if r + 1 < 4:
                    nb.append([i, i + 4])
        return nb
    raise ValueError(lattice)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while running."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_id):
        self.gpu_id = gpu_id
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_id), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, busy = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, cmax, util = float(parts[0]), float(parts[1]), float(parts[7])
            except ValueError:
                continue
            sm.append(clk); smax.append(cmax)
            if util >= 50:
                busy.append(clk)
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        use = busy if busy else sm
        use = sorted(use)
        med = use[len(use) // 2] if use else None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy)}


# ---------------------------------------------------------------------------------------
# reference arm: CPU port of the reference H.v on a bounded sample
# ---------------------------------------------------------------------------------------

def host_threads():
    """Host threads the CPU arm uses: every core this process may run on.  Set explicitly --
    torchrun exports OMP_NUM_THREADS=1 to its children, which must not shrink the baseline."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_port(workload):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c

    num_sites, lattice, n_up, n_dn = WORKLOADS[workload]
    nb = neighbors_of(lattice, num_sites)
    return oracle_c.hubbard_oracle(num_sites, n_up, n_dn, nb, PARAMS["inter"], -PARAMS["mu"], PARAMS["hop"])


def cpu_port_rate(workload, target_seconds, steps, warmup, orc=None):
    """Times oracle/hv_oracle.c (OpenMP over all host threads) on the workload.  The whole sector
    is applied whenever `steps + warmup` full H.v fit `target_seconds`; otherwise a contiguous block
    of up-rows from the middle of the sector, extrapolated linearly (said in `sample`)."""
    import numpy as np

    orc = build_port(workload) if orc is None else orc
    num_up = len(orc.up)
    x = np.random.default_rng(0).standard_normal(orc.size)
    x /= np.linalg.norm(x)
    cores = host_threads()
    probe = min(num_up, max(cores * 4, 64))
    row0 = (num_up - probe) // 2
    orc.matvec_rows(x, row0, probe, nthreads=cores)          # first touch / thread pool start
    t0 = time.perf_counter()
    orc.matvec_rows(x, row0, probe, nthreads=cores)
    per_row = (time.perf_counter() - t0) / probe
    total_steps = max(1, steps + warmup)
    nrows = int(min(num_up, max(probe, target_seconds / total_steps / max(per_row, 1e-9))))
    row0 = (num_up - nrows) // 2
    for _ in range(warmup):
        orc.matvec_rows(x, row0, nrows, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.matvec_rows(x, row0, nrows, nthreads=cores)
    dt = (time.perf_counter() - t0) / steps
    full = dt * num_up / nrows
    what = "the whole sector" if nrows == num_up else "middle of the sector, extrapolated linearly"
    info = {"cores": cores, "kind": "port",
            "sample": f"{nrows} of {num_up} up-rows ({nrows * len(orc.dn)} of {orc.size} states) per "
                      f"step, {what}; oracle/hv_oracle.c, OpenMP with {cores} threads set explicitly"}
    return 1.0 / full, dt * 1e3, info


def cpu_port_lanczos(workload, orc=None, budget_s=120.0):
    """Second half of the metric on the CPU arm: the reference's ground-state call
    `sla.eigsh(hamop, k=1, which="SA")` (cmpy/exactdiag.py:37) with the C port as the operator's
    mat-vec, tol 1e-10.  Bounded: the mat-vec raises once `budget_s` is used up and the record then
    says how far ARPACK got (the run must stay within minutes)."""
    import numpy as np
    import scipy.sparse.linalg as sla

    orc = build_port(workload) if orc is None else orc
    cores = host_threads()
    n = orc.size
    x = np.random.default_rng(0).standard_normal(n)
    x /= np.linalg.norm(x)
    ncv = 12
    count = [0]
    t_start = time.perf_counter()

    class _Budget(Exception):
        pass

    def mv(v):
        if time.perf_counter() - t_start > budget_s:
            raise _Budget()
        count[0] += 1
        return orc.matvec(np.ascontiguousarray(v, dtype=np.float64).reshape(-1), nthreads=cores)

    op = sla.LinearOperator((n, n), matvec=mv, dtype=np.float64)
    base = {"tol": 1e-10, "ncv": ncv, "cores": cores, "dim": n,
            "solver": "scipy.sparse.linalg.eigsh(k=1, which='SA') over the C port (cmpy/exactdiag.py:37)"}
    try:
        ev = sla.eigsh(op, k=1, which="SA", tol=1e-10, ncv=ncv, v0=x, return_eigenvectors=False)
    except _Budget:
        return dict(base, seconds=None, matvecs=count[0], elapsed_s=time.perf_counter() - t_start,
                    skipped=f"not converged within the {budget_s:.0f} s budget of the bounded CPU arm")
    return dict(base, seconds=time.perf_counter() - t_start, e0=float(ev[0]), matvecs=count[0])


def cpu_scipy_path(max_sites=12):
    """BASELINE.md section 3 baseline B -- the reference's scipy CSR path (`hubbard_hamiltonian`,
    cmpy/models/hubbard.py:25-34: COO triplets -> csr_matrix -> A @ x, eigsh(A, k=1, which='SA')) --
    restated with oracle/oracle_np.py (kind "port": /root/reference does not exist on the GPU box),
    open chains at half filling, single-threaded SpMV as in the reference."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    import oracle_np as orc

    out = []
    for L in (8, 10, 12):
        if L > max_sites:
            break
        n = L // 2
        st = orc.enumerate_states(L, n)
        t0 = time.perf_counter()
        r, c, v = orc.hubbard_triplets(st, st, L, orc.chain_neighbors(L), PARAMS["inter"], -PARAMS["mu"],
                                       PARAMS["hop"])
        dim = len(st) ** 2
        a = sp.csr_matrix((v, (r, c)), shape=(dim, dim))
        t_build = time.perf_counter() - t0
        x = np.random.default_rng(0).standard_normal(dim)
        reps = max(3, int(2e7 / max(a.nnz, 1)))
        a @ x
        t0 = time.perf_counter()
        for _ in range(reps):
            a @ x
        t_mv = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        ev = sla.eigsh(a, k=1, which="SA", tol=1e-10, return_eigenvectors=False)
        t_e0 = time.perf_counter() - t0
        out.append({"config": f"hubbard_chain_L{L}_half_filling", "dim": dim, "nnz": int(a.nnz),
                    "build_s": t_build, "matvec_per_s": 1.0 / t_mv, "eigsh_e0_s": t_e0, "e0": float(ev[0])})
    return out


def reference_verbatim_record():
    """BASELINE.md section 3 baseline A (the unmodified reference: hamilton_operator +
    HamiltonOperator.matvec + sla.eigsh) cannot run on the GPU box; the numbers measured in the
    dev container by tools/reference_cpu_baseline.py are committed and quoted here."""
    path = os.path.join(ROOT, "profiles", "r2_reference_cpu_baseline.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = build_port(args.workload)
    rate, ms_sample, info = cpu_port_rate(args.workload, 60.0, args.steps, args.warmup, orc)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": dict(value=rate, unit=UNIT, **info),
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_ms_per_step": ms_sample,
    }
    if not args.no_lanczos:
        try:   # a size the CPU arm always finishes: same call, 14-site chain (dim 11 778 624)
            line["lanczos_e0_chain14"] = cpu_port_lanczos("chain14", None, 120.0)
        except Exception as exc:  # never in the way of the contract line
            line["lanczos_e0_chain14"] = {"skipped": repr(exc)}
        try:
            line["lanczos_e0"] = cpu_port_lanczos(args.workload, orc)
        except Exception as exc:
            line["lanczos_e0"] = {"skipped": repr(exc)}
    print(json.dumps(line), flush=True)


def workload_config(workload, n_gpus):
    num_sites, lattice, n_up, n_dn = WORKLOADS[workload]
    from math import comb

    dim = comb(num_sites, n_up) * comb(num_sites, n_dn)
    return {
        "workload": f"hubbard_{lattice}_L{num_sites}_nup{n_up}_ndn{n_dn}_hv",
        "baseline_config": {"c4": "configs[3] Hubbard 4x4 half filling", "c2": "configs[1]",
                            "c1": "configs[0]"}.get(workload, workload),
        "dim": dim, "U": PARAMS["inter"], "mu": PARAMS["mu"], "hop": PARAMS["hop"],
        "bytes_per_step_algorithmic": 16 * dim,
        "l2_policy": ("inputs larger than L2 (8*dim bytes per vector), no flush"
                      if 8 * dim > 200e6 else "L2 flushed between timed iterations"),
    }


def l2_to_sm_roofline(workload, ms_per_step):
    """Secondary roofline of the single-pass H.v: besides its own row every
    amplitude pulls `hops` up-hop neighbours through L2 -> SM, hops = sum over bonds of the probability
    that the two sites are occupied differently = nbonds * 2 n (L - n) / (L (L - 1)); the ceiling is
    the gather-only rate measured with tools/microbench.cu (profiles/r1_microbench_4x4.txt)."""
    from math import comb

    num_sites, lattice, n_up, n_dn = WORKLOADS[workload]
    nbonds = len(neighbors_of(lattice, num_sites))
    hops = nbonds * 2.0 * n_up * (num_sites - n_up) / (num_sites * (num_sites - 1))
    dim = comb(num_sites, n_up) * comb(num_sites, n_dn)
    nbytes = (hops + 1.0) * 8.0 * dim
    achieved = nbytes / (ms_per_step * 1e-3) / 1e9
    peak = 8700.0
    return {"bound": "l2_to_sm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "bytes_per_launch": nbytes, "up_hops_per_state": hops,
            "peak_source": "measured gather-only kernel, profiles/r1_microbench_4x4.txt"}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------

def parity_check(workload, x, y, world, dev, hamop):
    """Outside the timed region: rows of y = H x of the benchmarked operator against the C oracle
    (oracle/hv_oracle.c).  N = 1: 32-row blocks at the start, middle and end of the sector; N > 1:
    the same for rank 0's slab (x of the other ranks is regenerated from their seeds)."""
    import numpy as np
    import torch

    orc = build_port(workload)
    nd = len(orc.dn)
    if world == 1:
        xf = x.cpu().numpy()
        r_lo, r_hi = 0, len(orc.up)
    else:   # (y was computed collectively by the caller; only rank 0 gets here)
        parts = []
        for r in range(world):
            a, b = hamop.plan.rows(r)
            g = torch.Generator(device=dev); g.manual_seed(r)
            parts.append(torch.randn((b - a) * nd, dtype=torch.float64, device=dev, generator=g).cpu().numpy())
        xf = np.concatenate(parts)
        r_lo, r_hi = hamop.plan.rows(0)
    nblk = min(32, r_hi - r_lo)
    worst, scale = 0.0, 0.0
    rows = sorted({r_lo, (r_lo + r_hi - nblk) // 2, r_hi - nblk})
    for r0 in rows:
        ref = orc.matvec_rows(xf, r0, nblk, nthreads=host_threads())
        got = y[(r0 - r_lo) * nd:(r0 - r_lo + nblk) * nd].cpu().numpy()
        scale = max(scale, float(np.abs(ref).max()))
        worst = max(worst, float(np.abs(got - ref).max()))
    return {"max_rel_err": worst / max(scale, 1e-300), "rows_checked": [int(r) for r in rows],
            "rows_per_block": int(nblk), "oracle": "oracle/hv_oracle.c", "tolerance": 1e-12}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from cmpy_b200 import _lib
    from cmpy_b200.models import HubbardModel
    from cmpy_b200.exactdiag import lanczos_run

    num_sites, lattice, n_up, n_dn = WORKLOADS[args.workload]
    nb = neighbors_of(lattice, num_sites)
    model = HubbardModel(num_sites, nb, **PARAMS)
    dev = torch.device("cuda", local_rank)
    flush = None

    if world == 1:
        hamop = model.hamilton_operator(n_up, n_dn)
        dim = hamop.shape[0]
        g = torch.Generator(device=dev); g.manual_seed(0)
        x = torch.randn(dim, dtype=torch.float64, device=dev, generator=g)
        x /= x.norm()
        y = torch.empty_like(x)
        if 8 * dim <= 200e6:
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

        def step():
            hamop.apply(x, out=y)

        local_elems = dim
    else:
        from cmpy_b200.dist import ShardedHubbardOperator

        hamop = ShardedHubbardOperator(model, n_up, n_dn)
        dim = hamop.shape[0]
        g = torch.Generator(device=dev); g.manual_seed(rank)
        x = torch.randn(hamop.local_size, dtype=torch.float64, device=dev, generator=g)
        y = torch.empty_like(x)

        def step():
            hamop.apply_local(x, out=y)

        local_elems = hamop.local_size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_buf=None):
        """K steps bracketed by barrier+sync, CUDA events on the launching stream, max over
        ranks. With an L2 flush between iterations the flush is excluded (per-step events)."""
        barrier()
        if flush_buf is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(steps):
                flush_buf.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = None
    if rank == 0:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        sampler = ClockSampler("GPU-" + uuid if not uuid.startswith("GPU-") else uuid)
        sampler.start()

    for _ in range(max(args.warmup, 3)):
        step()
    _lib.reset_launch_count()
    ms_total = timed(step, args.steps, flush)
    launches = _lib.launch_count()
    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step

    # ---- e2e: pinned host buffers, H2D + H.v + D2H per step -------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    xh = torch.empty(local_elems, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    h2d = d2h = 8 * local_elems * world

    yh_out = torch.empty(local_elems, dtype=torch.float64).pin_memory()

    def e2e_step():
        return hamop.matvec(xh, out=yh_out)  # pinned CPU tensor in -> pinned CPU tensor out

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_single = e2e_steps / e2e_s   # one blocking call per step: H2D, H.v, D2H back to back
    # the batched public call (column-by-column matmat of the reference on host data), the SAME API at
    # every N: every step still copies its own input from pinned host memory and its result back, but
    # the two PCIe directions and the kernel of consecutive steps overlap
    hamop.matvec_batch([xh, xh])
    barrier()
    t0 = time.perf_counter()
    hamop.matvec_batch([xh] * e2e_steps)
    barrier()
    e2e_b = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_b], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_b = float(t.item())
    e2e_value = e2e_steps / e2e_b

    # the call scipy makes (eigsh / expm_multiply hand `_matvec` a pageable numpy vector)
    e2e_numpy = None
    if world == 1:
        xn = xh.numpy().copy()
        hamop.matvec(xn)
        t0 = time.perf_counter()
        for _ in range(3):
            hamop.matvec(xn)
        e2e_numpy = 3 / (time.perf_counter() - t0)
        del xn

    # ---- parity of THIS configuration against the C oracle, outside the timed region ---------
    parity = None
    if not args.no_parity:
        step()                     # collective on every rank: y = (H x)_local of the timed configuration
        torch.cuda.synchronize()
        if rank == 0:
            try:
                parity = parity_check(args.workload, x, y, world, dev, hamop)
            except Exception as exc:
                parity = {"error": repr(exc)}
    phases = None
    if world > 1 and getattr(hamop, "exchange", "") == "peer":
        try:
            phases = hamop.profile_phases(x, y, reps=5)
        except Exception as exc:
            phases = {"error": repr(exc)}
        step()

    # ---- Lanczos E0 to 1e-10 (second half of the BASELINE metric), single GPU only --------
    lanczos = None
    if world == 1 and not args.no_lanczos:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = lanczos_run(hamop, None, maxit=1000, tol=1e-10, check_every=10)
        torch.cuda.synchronize()
        lanczos = {"seconds": time.perf_counter() - t0, "iterations": res.nit, "e0": res.e0,
                   "converged": bool(res.converged), "tol": 1e-10,
                   "bytes_per_iteration_algorithmic": 48 * dim}

        if args.workload != "chain14":   # the size the CPU arm's eigsh leg always reaches
            ns14, lat14, nu14, nd14 = WORKLOADS["chain14"]
            h14 = HubbardModel(ns14, neighbors_of(lat14, ns14), **PARAMS).hamilton_operator(nu14, nd14)
            lanczos_run(h14, None, maxit=20, tol=1e-10, check_every=10)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r14 = lanczos_run(h14, None, maxit=1000, tol=1e-10, check_every=10)
            torch.cuda.synchronize()
            lanczos["chain14"] = {"seconds": time.perf_counter() - t0, "iterations": r14.nit, "e0": r14.e0,
                                  "converged": bool(r14.converged), "dim": h14.shape[0]}
            del h14

    if world > 1 and not args.no_lanczos:
        from cmpy_b200.dist import lanczos_sharded

        lanczos_sharded(hamop, maxit=10, tol=1e-10, check_every=10)      # warm-up
        barrier()
        t0 = time.perf_counter()
        e0s, _, _, nits, convs = lanczos_sharded(hamop, maxit=1000, tol=1e-10, check_every=10)
        barrier()
        lanczos = {"seconds": time.perf_counter() - t0, "iterations": int(nits), "e0": float(e0s),
                   "converged": bool(convs), "tol": 1e-10,
                   "path": getattr(hamop, "last_lanczos_path", "python")}

    clocks = sampler.stop() if sampler is not None else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = 16.0 * dim / world / (ms_per_step * 1e-3) / 1e9  # per GPU
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "hv_traffic.json")
        if world == 1 and os.path.exists(tpath):   # ncu figure of the single-GPU kernel only
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "achieved_hbm_gbs_algorithmic": achieved * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("hub_seg_kernel<UNI, LZ=0, VEC2, 896> (default full H.v)"
                                    if world == 1 else "sharded step (per GPU): dn pass, push, up pass, pull"),
                         "algorithmic_bytes_per_launch": 16 * dim // world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": ("HamiltonOperator.matvec_batch (pipelined host batch)" if world == 1
                            else "ShardedHubbardOperator.matvec_batch (pipelined host batch of the local slabs)"),
                    "single_call_value": e2e_single},
            "gpu_launches": launches,
            "clocks": clocks,
            "parallelism": ("single GPU" if world == 1 else
                            f"up-string sharded x{world}; " + (
                                "peer-memory push/pull transposes over NVLink (symmetric memory), NCCL for "
                                "barriers / scalar all-reduces only" if getattr(hamop, "exchange", "") == "peer"
                                else "NCCL all-to-all transposes")),
        }
        if parity is not None:
            line["parity_max_rel_err"] = parity.get("max_rel_err")
            line["parity"] = parity
        if e2e_numpy is not None:
            line["e2e"]["numpy_call_value"] = e2e_numpy
            line["e2e"]["numpy_call_api"] = "HamiltonOperator.matvec(np.ndarray) -- the call sla.eigsh makes"
        if world > 1:
            out_bytes = hamop.plan.bytes_out_per_hv()
            nv = out_bytes / (ms_per_step * 1e-3) / 1e9
            line["roofline_nvlink"] = {"bound": "nvlink", "bytes_out_per_gpu_per_step": out_bytes,
                                       "achieved": nv, "peak": 900.0, "unit": "GB/s", "frac": nv / 900.0,
                                       "frac_of_measured_peer_copy_770": nv / 770.0,
                                       "note": "bytes leaving rank 0 per H.v (two transposes) / whole step time"}
            if phases is not None:
                line["phases_ms_serialised"] = phases
        if world == 1:
            try:
                line["roofline_l2_to_sm"] = l2_to_sm_roofline(args.workload, ms_per_step)
            except Exception:  # explanatory extra, never in the way of the contract line
                pass
        if lanczos is not None:
            line["lanczos_e0"] = lanczos
        if world == 1 and not args.no_cpu_baseline:
            rate, ms_sample, info = cpu_port_rate(args.workload, 12.0, 2, 1)
            line["cpu_baseline"] = dict(value=rate, unit=UNIT, **info)
            extra = {}
            try:
                extra["scipy_csr_path_port"] = cpu_scipy_path()
            except Exception as exc:
                extra["scipy_csr_path_port"] = {"error": repr(exc)}
            rec = reference_verbatim_record()
            if rec is not None:
                extra["reference_verbatim_dev_container"] = rec
            line["cpu_baseline"]["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-lanczos", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
