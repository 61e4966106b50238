# -*- coding: utf-8 -*-
"""DEV-CONTAINER ONLY (needs /root/reference): BASELINE.md section 3 baseline A -- the unmodified
reference timed on this container's CPU: `model.hamilton_operator(...)`, `HamiltonOperator.matvec`
(cmpy/operators.py:626-630), `sla.eigsh(hamop, k=1, which="SA")` (cmpy/exactdiag.py:37),
`gf_lehmann` (cmpy/exactdiag.py:215-245) and the Heisenberg operator (cmpy/models/heisenberg.py:19-40),
plus baseline B through the reference's own `hubbard_hamiltonian` (cmpy/models/hubbard.py:25-34).
Writes profiles/r2_reference_cpu_baseline.json, which bench.py quotes under `cpu_baseline.extra`
(the GPU box has no /root/reference).  Run: python tools/reference_cpu_baseline.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refshim  # noqa: E402

cmpy = refshim.load_reference()
import scipy.sparse.linalg as sla  # noqa: E402
from cmpy.models import HubbardModel  # noqa: E402
from cmpy.models.hubbard import hubbard_hamiltonian  # noqa: E402
from cmpy.exactdiag import gf_lehmann  # noqa: E402


def timeit(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


out = {"where": "dev container CPU, single-threaded Python + numba JIT as shipped by the reference",
       "cpu_count": os.cpu_count(), "hubbard_chain_half_filling": [], "heisenberg_chain_sz0": []}
for L in (6, 8, 10):
    nb = [[i, i + 1] for i in range(L - 1)]
    model = HubbardModel(L, nb, inter=4.0, mu=2.0, hop=1.0)
    t0 = time.perf_counter()
    sector = model.get_sector(L // 2, L // 2)
    t_sector = time.perf_counter() - t0
    t0 = time.perf_counter()
    hamop = model.hamilton_operator(sector=sector)
    t_build = time.perf_counter() - t0
    dim = hamop.shape[0]
    x = np.random.default_rng(0).standard_normal(dim)
    t_mv = timeit(lambda: hamop.matvec(x), 3 if L < 10 else 1)
    t0 = time.perf_counter()
    ev = sla.eigsh(hamop, k=1, which="SA", tol=1e-10, return_eigenvectors=False)
    t_e0 = time.perf_counter() - t0
    # baseline B: the reference's scipy CSR helper
    t0 = time.perf_counter()
    a = hubbard_hamiltonian(sector, nb, inter=4.0, eps=-2.0, hop=1.0)
    t_csr = time.perf_counter() - t0
    rec = {"L": L, "dim": dim, "sector_s": t_sector, "operator_build_s": t_build,
           "A_matvec_per_s": 1.0 / t_mv, "A_eigsh_e0_s": t_e0, "e0": float(ev[0])}
    if a is not None:
        t_spmv = timeit(lambda: a @ x, 20)
        t0 = time.perf_counter()
        ev2 = sla.eigsh(a, k=1, which="SA", tol=1e-10, return_eigenvectors=False)
        rec.update({"B_csr_build_s": t_csr, "B_matvec_per_s": 1.0 / t_spmv,
                    "B_eigsh_e0_s": time.perf_counter() - t0, "B_e0": float(ev2[0])})
    if L <= 6:
        z = np.linspace(-6, 6, 1001) + 0.05j
        t0 = time.perf_counter()
        gf_lehmann(model, z, beta=10.0, pos=0)
        rec["gf_lehmann_1001_s"] = time.perf_counter() - t0
    out["hubbard_chain_half_filling"].append(rec)
    print(rec, flush=True)

from cmpy.models import HeisenbergModel  # noqa: E402

for N in (8, 10, 12):
    latt = refshim.ChainStandIn(N)
    model = HeisenbergModel(latt, j=1.0, jz=1.0)
    t0 = time.perf_counter()
    hamop = model.hamilton_operator(s=0)
    t_build = time.perf_counter() - t0
    dim = hamop.shape[0]
    x = np.random.default_rng(0).standard_normal(dim)
    t_mv = timeit(lambda: hamop.matvec(x), 3)
    t0 = time.perf_counter()
    ev = sla.eigsh(hamop, k=1, which="SA", tol=1e-10, return_eigenvectors=False)
    rec = {"N": N, "dim": dim, "operator_build_s": t_build, "matvec_per_s": 1.0 / t_mv,
           "eigsh_e0_s": time.perf_counter() - t0, "e0": float(ev[0])}
    out["heisenberg_chain_sz0"].append(rec)
    print(rec, flush=True)

path = os.path.join(ROOT, "profiles", "r2_reference_cpu_baseline.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
