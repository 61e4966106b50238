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

def cpu_port_rate(workload, target_seconds, steps, warmup):
    """Times oracle/hv_oracle.c (OpenMP, all host threads) on a contiguous sample of
    up-rows of the workload; returns (matvec/s extrapolated to the full sector, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_c

    num_sites, lattice, n_up, n_dn = WORKLOADS[workload]
    nb = neighbors_of(lattice, num_sites)
    orc = oracle_c.hubbard_oracle(num_sites, n_up, n_dn, nb, PARAMS["inter"], -PARAMS["mu"],
                                  PARAMS["hop"])
    num_up = len(orc.up)
    x = np.random.default_rng(0).standard_normal(orc.size)
    x /= np.linalg.norm(x)
    cores = oracle_c.max_threads()
    # calibrate the sample size on a few rows (spread over the sector so gathers are typical)
    probe = min(num_up, max(cores * 2, 16))
    row0 = (num_up - probe) // 2
    t0 = time.perf_counter()
    orc.matvec_rows(x, row0, probe)
    t_probe = time.perf_counter() - t0
    per_row = t_probe / probe
    total_steps = max(1, steps + warmup)
    nrows = int(min(num_up, max(probe, target_seconds / total_steps / max(per_row, 1e-9))))
    row0 = (num_up - nrows) // 2
    for _ in range(warmup):
        orc.matvec_rows(x, row0, nrows)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.matvec_rows(x, row0, nrows)
    dt = (time.perf_counter() - t0) / steps
    full = dt * num_up / nrows
    info = {"cores": cores, "kind": "port",
            "sample": f"{nrows} of {num_up} up-rows ({nrows * len(orc.dn)} of {orc.size} states) per "
                      f"step, middle of the sector, extrapolated linearly; oracle/hv_oracle.c OpenMP"}
    return 1.0 / full, dt * 1e3, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms_sample, info = cpu_port_rate(args.workload, 60.0, args.steps, args.warmup)
    num_sites, lattice, n_up, n_dn = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, 1),
        "cpu_baseline": dict(value=rate, unit=UNIT, **info),
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_ms_per_step": ms_sample,
    }
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
        "parallelism": "single GPU" if n_gpus == 1 else f"up-string sharded x{n_gpus}, NCCL all-to-all",
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

    def e2e_step():
        yh = hamop.matvec(xh)  # CPU (pinned) tensor in -> pinned CPU tensor out
        return yh

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
    e2e_value = e2e_single
    if world == 1:
        # the batched public call (column-by-column matmat of the reference on host data): every
        # step still copies its own input from pinned host memory and its result back, but the
        # two PCIe directions and the kernel of consecutive steps overlap
        hamop.matvec_batch([xh, xh])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hamop.matvec_batch([xh] * e2e_steps)
        torch.cuda.synchronize()
        e2e_value = e2e_steps / (time.perf_counter() - t0)

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

    clocks = sampler.stop() if sampler is not None else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = 16.0 * dim / world / (ms_per_step * 1e-3) / 1e9  # per GPU
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "hv_traffic.json")
        if os.path.exists(tpath):
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
                         "kernel": "hub_seg_kernel" if world == 1 else "sharded step (per GPU)",
                         "algorithmic_bytes_per_launch": 16 * dim // world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": ("HamiltonOperator.matvec_batch (pipelined host batch)" if world == 1
                            else "ShardedHubbardOperator.matvec (blocking call per step)"),
                    "single_call_value": e2e_single},
            "gpu_launches": launches,
            "clocks": clocks,
        }
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
