"""CPU emulation of the generation-3 row engine (cmpy_b200/csrc/hubbard_eng.cuh): the staging map and
the two shared-memory phase bodies (__host__ __device__) run lane by lane from the library's own table
builder and are compared with a direct evaluation of one row of (D + T_dn) x
(ref: cmpy/operators.py:305-527).  Test infrastructure only (tests/emu/eng_emu.cu)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_np as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "eng_emu.cu")
HDR = os.path.join(os.path.dirname(HERE), "cmpy_b200", "csrc", "hubbard_eng.cuh")
OUT = os.path.join(HERE, "emu", "_build", "libeng_emu.so")


def _build():
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        return OUT
    if shutil.which("nvcc") is None:
        if os.path.exists(OUT):
            return OUT
        pytest.skip("nvcc not available to build the emulation harness")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                    "-Xcompiler", "-fPIC", "-o", OUT, SRC], check=True)
    return OUT


@pytest.fixture(scope="module")
def emu():
    lib = ctypes.CDLL(_build())
    lib.eng_emu_row.restype = ctypes.c_int
    lib.eng_emu_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, ctypes.c_uint, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_int)]
    return lib


def ring(n):
    return orc.chain_neighbors(n, True)


def ladder(n):
    nb = [[i, i + 1] for i in range(0, n // 2 - 1)] + [[i, i + 1] for i in range(n // 2, n - 1)]
    return nb + [[i, i + n // 2] for i in range(n // 2)]


def direct_row(L, n_dn, bonds, width, eps0, u0, hop0, ups, e_up, x):
    dn = orc.enumerate_states(L, n_dn)
    e_dn = 0.0
    for _ in range(n_dn):
        e_dn += eps0
    diag = e_up + e_dn + u0 * np.array([int(ups & int(s)).bit_count() for s in dn], dtype=np.float64)
    y = diag * x
    for i, j in bonds:
        org, tgt, sgn = orc.species_hops(dn, width, i, j)
        y[org] += hop0 * sgn * x[tgt]
    return y


CASES = [
    (16, 8, "sq44", 16), (16, 8, "chain", 16), (16, 7, "sq44", 16), (16, 3, "ring", 16), (16, 13, "sq44", 16),
    (16, 8, "sq44", 0), (15, 7, "chain", 15), (14, 7, "ring", 14), (13, 6, "chain", 13), (12, 6, "sq43", 12),
    (12, 5, "ladder", 12), (10, 5, "ring", 10), (8, 4, "chain", 8), (8, 1, "ring", 8), (8, 7, "chain", 8),
    (6, 3, "ring", 6), (5, 2, "chain", 5), (4, 2, "sq22", 4), (16, 1, "sq44", 16), (16, 15, "chain", 16),
    (16, 0, "chain", 16), (16, 16, "chain", 16), (11, 4, "ring", 11), (9, 4, "sq33", 9),
    # dense halves: lists of more than four entries of one sign (the walk beyond the packed descriptor words)
    (16, 8, "dense", 16), (12, 6, "dense", 12), (14, 5, "dense", 0),
]


def dense_halves(L):
    """All pairs inside the low half, all pairs inside the high half, two bonds across."""
    m = (L + 1) // 2
    b = [(i, j) for i in range(m) for j in range(i + 1, m)] + [(i, j) for i in range(m, L) for j in range(i + 1, L)]
    return b + [(m - 1, m), (0, L - 1)]


@pytest.mark.parametrize("L,n_dn,lat,width", CASES)
@pytest.mark.parametrize("nwarps", [32, 16])
def test_engine_row_matches_direct(emu, L, n_dn, lat, width, nwarps):
    bonds = {"chain": lambda: orc.chain_neighbors(L), "ring": lambda: ring(L), "ladder": lambda: ladder(L),
             "sq44": lambda: orc.square_neighbors(4, 4), "sq43": lambda: orc.square_neighbors(4, 3),
             "sq33": lambda: orc.square_neighbors(3, 3), "sq22": lambda: orc.square_neighbors(2, 2),
             "dense": lambda: dense_halves(L)}[lat]()
    bonds = sorted({(min(i, j), max(i, j)) for i, j in bonds if i != j})
    s1 = (ctypes.c_int * len(bonds))(*[b[0] for b in bonds])
    s2 = (ctypes.c_int * len(bonds))(*[b[1] for b in bonds])
    from math import comb

    nd = comb(L, n_dn)
    rng = np.random.default_rng(L * 100 + n_dn)
    for trial in range(2):
        x = rng.standard_normal(nd)
        ups = int(rng.integers(0, 1 << L))
        e_up, eps0, u0, hop0 = 0.37 * trial - 1.1, -2.0 + trial, 4.0 - 1.5 * trial, 1.0 if trial == 0 else -0.8
        y = np.full(nd, np.nan)
        info = (ctypes.c_int * 8)()
        rc = emu.eng_emu_row(L, n_dn, len(bonds), s1, s2, width, eps0, u0, hop0, ups, e_up, nwarps, 0,
                             x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                             y.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), info)
        if rc == 1:
            pytest.skip("sector outside the engine's range")
        assert rc == 0, rc
        ref = direct_row(L, n_dn, bonds, width, eps0, u0, hop0, ups, e_up, x)
        assert np.isfinite(y).all()
        assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
        assert info[4] <= 32000   # the constant-bank table must fit the kernel-parameter space


def test_engine_random_graphs(emu):
    """Random bond graphs (dense or sparse inside the halves, at most four bonds across), fillings, sign widths and
    warp counts: the descriptor tables (packed list entries, LH ranks / flags) against the direct evaluation."""
    from math import comb

    rng = np.random.default_rng(777)
    ran = 0
    for trial in range(40):
        L = int(rng.integers(6, 17))
        m = (L + 1) // 2
        lo = [(i, j) for i in range(m) for j in range(i + 1, m)]
        hi = [(i, j) for i in range(m, L) for j in range(i + 1, L)]
        cross = [(i, j) for i in range(m) for j in range(m, L)]
        pick = lambda pool, k: [pool[t] for t in rng.choice(len(pool), size=min(k, len(pool)), replace=False)] if pool else []
        bonds = sorted(set(pick(lo, int(rng.integers(0, 12))) + pick(hi, int(rng.integers(0, 12))) + pick(cross, int(rng.integers(0, 5)))))
        if not bonds:
            continue
        n_dn = int(rng.integers(1, L))
        width = L if trial % 4 else 0
        nwarps = int(rng.choice([32, 16, 7]))
        s1 = (ctypes.c_int * len(bonds))(*[b[0] for b in bonds])
        s2 = (ctypes.c_int * len(bonds))(*[b[1] for b in bonds])
        nd = comb(L, n_dn)
        x = rng.standard_normal(nd)
        ups = int(rng.integers(0, 1 << L))
        e_up, eps0, u0, hop0 = float(rng.normal()), float(rng.normal()), float(rng.normal()), float(rng.choice([1.0, -0.7, 2.5]))
        y = np.full(nd, np.nan)
        info = (ctypes.c_int * 8)()
        rc = emu.eng_emu_row(L, n_dn, len(bonds), s1, s2, width, eps0, u0, hop0, ups, e_up, nwarps, 0,
                             x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                             y.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), info)
        if rc == 1:
            continue          # outside the engine's range (row too short, list too long ...)
        assert rc == 0, (rc, L, bonds, n_dn)
        ref = direct_row(L, n_dn, bonds, width, eps0, u0, hop0, ups, e_up, x)
        assert np.isfinite(y).all()
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (L, bonds, n_dn, width, nwarps)
        ran += 1
    assert ran >= 15, ran
