# -*- coding: utf-8 -*-
"""CPU emulation of the shared-memory phases of the class-major H.v kernel
(cmpy_b200/csrc/hubbard_cls.cuh): the __host__ __device__ phase bodies of engine 0 (the one
measured on B200) are run lane by lane by tests/emu/cls_emu.cu and
compared with a direct evaluation of (D + T_dn) x on one row of the amplitude matrix
(matrix elements: cmpy/operators.py:305-527; Heisenberg flavour: cmpy/models/heisenberg.py:19-40).
This checks the table construction and the index arithmetic of the phases without a GPU; the
GPU parity tests (tests/test_gpu_parity.py) check the kernels themselves."""
import ctypes
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_np as orc  # noqa: E402

SRC = os.path.join(HERE, "emu", "cls_emu.cu")
OUT = os.path.join(HERE, "emu", "_build", "libcls_emu.so")


def _build():
    import shutil

    if shutil.which("nvcc") is None:
        if os.path.exists(OUT):
            return OUT   # prebuilt by __graft_entry__.build()
        pytest.skip("nvcc not available to build the emulation harness")
    deps = [SRC] + [os.path.join(ROOT, "cmpy_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "cmpy_b200", "csrc"))
                    if f.endswith(".cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC", "-o", OUT, SRC], check=True)
    return OUT


@pytest.fixture(scope="module")
def emu():
    lib = ctypes.CDLL(_build())
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    lib.emu_cls_row.restype = ctypes.c_int
    lib.emu_cls_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_double,
                                ctypes.c_double, ctypes.c_double, ctypes.c_uint, ctypes.c_double, ctypes.c_int,
                                ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, dp, dp, ip]
    return lib


def run_emu(lib, L, n_dn, bonds, width, u0, hop0, ups, eu, eng, x, spin=False, sd=(0.0, 0.0), nwarps=32):
    s1 = np.ascontiguousarray([b[0] for b in bonds], dtype=np.int32)
    s2 = np.ascontiguousarray([b[1] for b in bonds], dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    info = np.zeros(16, dtype=np.int32)
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    rc = lib.emu_cls_row(L, n_dn, len(bonds), s1.ctypes.data_as(ip), s2.ctypes.data_as(ip), width, 0.0, u0, hop0,
                         int(ups), eu, eng, int(spin), sd[0], sd[1], nwarps, x.ctypes.data_as(dp),
                         y.ctypes.data_as(dp), info.ctypes.data_as(ip))
    return rc, y, info


def direct_row(L, n_dn, bonds, width, u0, hop0, ups, eu, x, spin=False, sd=(0.0, 0.0)):
    """(D + T_dn) x for one row, straight from the definitions."""
    dn = np.asarray(orc.enumerate_states(L, n_dn), dtype=np.int64)
    rank = {int(s): i for i, s in enumerate(dn)}
    y = np.zeros_like(x)
    for i, s in enumerate(dn):
        s = int(s)
        if spin:
            anti = sum(1 for (a, b) in bonds if ((s >> a) & 1) != ((s >> b) & 1))
            diag = sd[0] + sd[1] * anti
        else:
            diag = eu + u0 * bin(ups & s).count("1")
        acc = 0.0
        for (a, b) in bonds:
            if ((s >> a) & 1) == ((s >> b) & 1):
                continue
            between = 0
            for k in range(a + 1, b):
                if k < width:
                    between |= 1 << k
            sign = -1.0 if bin(s & between).count("1") & 1 else 1.0
            acc += sign * x[rank[s ^ (1 << a) ^ (1 << b)]]
        y[i] = diag * x[i] + hop0 * acc
    return y


def chain(L):
    return [(i, i + 1) for i in range(L - 1)]


def ring(L):
    return chain(L) + [(0, L - 1)]


def square(nx, ny, periodic=False):
    b = []
    for r in range(ny):
        for c in range(nx):
            i = r * nx + c
            if c + 1 < nx:
                b.append((i, i + 1))
            elif periodic and nx > 2:
                b.append((r * nx, i))
            if r + 1 < ny:
                b.append((i, i + nx))
            elif periodic and ny > 2:
                b.append((c, i))
    return [(min(a, c), max(a, c)) for a, c in b]


CASES = [
    ("chain8", 8, 4, chain(8)),
    ("chain10", 10, 5, chain(10)),
    ("chain12", 12, 6, chain(12)),
    ("chain12_n4", 12, 4, chain(12)),
    ("ring12", 12, 6, ring(12)),
    ("sq4x3", 12, 6, square(4, 3)),
    ("chain14", 14, 7, chain(14)),
    ("chain16", 16, 8, chain(16)),
    ("sq4x4", 16, 8, square(4, 4)),
    ("chain16_n7_odd", 16, 6, chain(16)),
    ("ladder2x8", 16, 8, square(2, 8)),
]


@pytest.mark.parametrize("name,L,n_dn,bonds", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("eng", [0])
def test_hubbard_row(emu, name, L, n_dn, bonds, eng):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    num = len(orc.enumerate_states(L, n_dn))
    x = rng.standard_normal(num)
    ups = int(rng.integers(0, 1 << L))
    u0, hop0, eu = 4.0, 1.0, -3.25
    rc, y, info = run_emu(emu, L, n_dn, bonds, L, u0, hop0, ups, eu, eng, x)
    if rc == 1:
        pytest.skip("sector outside the class-major kernel (odd row length / too many straddling bonds)")
    assert rc == 0, rc
    ref = direct_row(L, n_dn, bonds, L, u0, hop0, ups, eu, x)
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (name, eng, np.abs(y - ref).max())


@pytest.mark.parametrize("eng", [0])
def test_c4_row_fits_and_matches(emu, eng):
    """4x4 sector (BASELINE config C4): the class-major row fits the shared-memory budget."""
    L, n = 16, 8
    rng = np.random.default_rng(5)
    x = rng.standard_normal(12870)
    rc, y, info = run_emu(emu, L, n, square(4, 4), L, 4.0, 1.0, 0b1010110010100110, -16.0, eng, x)
    assert rc == 0
    assert info[1] + 2560 <= 232448          # dynamic + static shared memory of one CTA
    ref = direct_row(L, n, square(4, 4), L, 4.0, 1.0, 0b1010110010100110, -16.0, x)
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("eng", [0])
@pytest.mark.parametrize("nwarps", [32, 24, 16, 5])
def test_task_distribution_independent_of_warp_count(emu, eng, nwarps):
    L, n = 12, 6
    rng = np.random.default_rng(7)
    x = rng.standard_normal(924)
    rc, y, _ = run_emu(emu, L, n, ring(L), L, 2.0, -0.7, 0b101100111000, 0.5, eng, x, nwarps=nwarps)
    assert rc == 0
    ref = direct_row(L, n, ring(L), L, 2.0, -0.7, 0b101100111000, 0.5, x)
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("eng", [0])
def test_signless_hops(emu, eng):
    """sign_width = 0 (Anderson convention, cmpy/models/anderson.py:149): no fermion signs."""
    L, n = 12, 6
    bonds = [(0, j) for j in range(1, L)]
    rng = np.random.default_rng(11)
    x = rng.standard_normal(924)
    rc, y, _ = run_emu(emu, L, n, bonds, 0, 0.0, 1.0, 0, 0.0, eng, x)
    if rc == 1:
        pytest.skip("star graph: more LH bonds than the class-major tables hold")
    assert rc == 0
    ref = direct_row(L, n, bonds, 0, 0.0, 1.0, 0, 0.0, x)
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("name,L,n,bonds", [("xxz_chain16", 16, 8, chain(16)), ("xxz_chain12", 12, 6, chain(12)),
                                             ("xxz_ladder", 16, 8, square(2, 8)), ("xxz_sub16_n9", 16, 9, chain(16))],
                         ids=["chain16", "chain12", "ladder2x8", "chain16_n9"])
@pytest.mark.parametrize("eng", [0])
def test_spin_flavour_row(emu, name, L, n, bonds, eng):
    """Heisenberg flavour: sign-free flips, Ising diagonal from antiparallel-bond counts."""
    rng = np.random.default_rng(3)
    num = len(orc.enumerate_states(L, n))
    x = rng.standard_normal(num)
    dz = 0.25
    sd = (dz * len(bonds), -2.0 * dz)
    rc, y, _ = run_emu(emu, L, n, bonds, 0, 0.0, 0.5, 0, 0.0, eng, x, spin=True, sd=sd)
    if rc == 1:
        pytest.skip("odd row length")
    assert rc == 0
    ref = direct_row(L, n, bonds, 0, 0.0, 0.5, 0, 0.0, x, spin=True, sd=sd)
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("eng", [0])
@pytest.mark.parametrize("L,n,bonds", [(16, 8, chain(16)), (16, 6, chain(16)), (12, 6, ring(12))], ids=["c16n8", "c16n6", "r12"])
def test_shifted_column_pairs(emu, eng, L, n, bonds):
    """Sub-rows of long rows that start at an odd element use the column pairs (2i-1, 2i) and the
    second pair table (pair_seg1): the slot map must still cover every column exactly once."""
    rng = np.random.default_rng(21)
    x = rng.standard_normal(len(orc.enumerate_states(L, n)))
    emu.emu_set_shift(1)
    try:
        rc, y, _ = run_emu(emu, L, n, bonds, L, 4.0, 1.0, 0b0110100110010110 & ((1 << L) - 1), -2.0, eng, x)
    finally:
        emu.emu_set_shift(0)
    assert rc == 0, rc
    ref = direct_row(L, n, bonds, L, 4.0, 1.0, 0b0110100110010110 & ((1 << L) - 1), -2.0, x)
    assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


def run_long(lib, L, n_dn, bonds, width, u0, hop0, ups, eu, eng, x):
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    lib.emu_long_row.restype = ctypes.c_int
    lib.emu_long_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_double,
                                 ctypes.c_double, ctypes.c_uint, ctypes.c_double, ctypes.c_int, dp, dp,
                                 ctypes.POINTER(ctypes.c_ubyte)]
    s1 = np.ascontiguousarray([b[0] for b in bonds], dtype=np.int32)
    s2 = np.ascontiguousarray([b[1] for b in bonds], dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    cov = np.zeros(len(x), dtype=np.uint8)
    rc = lib.emu_long_row(L, n_dn, len(bonds), s1.ctypes.data_as(ip), s2.ctypes.data_as(ip), width, u0, hop0,
                          int(ups), eu, eng, x.ctypes.data_as(dp), y.ctypes.data_as(dp),
                          cov.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    return rc, y, cov.astype(bool)


LONG_CASES = [
    ("chain17_n8", 17, 8, chain(17)),
    ("ring17_n5", 17, 5, ring(17)),
    ("lattice3x6_n9", 18, 9, square(6, 3)),
    ("chain19_n9_extra", 19, 9, chain(19) + [(3, 17), (12, 18), (16, 18)]),
    ("chain20_n4", 20, 4, chain(20)),
]


@pytest.mark.parametrize("name,L,n_dn,bonds", LONG_CASES, ids=[c[0] for c in LONG_CASES])
@pytest.mark.parametrize("eng", [0])
def test_long_row(emu, name, L, n_dn, bonds, eng):
    """Rows of more than 16 sites (BASELINE config C5 is the 20-site chain): the tables of
    build_long_tables (sub-rows by the top bits, top-bond gather lists, straddling-bond index maps)
    and the sub-row passes reproduce (D + T_dn) x on every column the class-major kernel takes."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    num = len(orc.enumerate_states(L, n_dn))
    x = rng.standard_normal(num)
    ups = int(rng.integers(0, 1 << L))
    u0, hop0, eu = 4.0, 1.0, -1.5
    rc, y, cov = run_long(emu, L, n_dn, bonds, L, u0, hop0, ups, eu, eng, x)
    if rc == 1:
        pytest.skip("no sub-row class of this sector fits the class-major kernel")
    assert rc == 0, rc
    assert cov.any()
    ref = direct_row(L, n_dn, bonds, L, u0, hop0, ups, eu, x)
    assert np.abs(y[cov] - ref[cov]).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name,N,n,bonds", [("xxz_chain18", 18, 9, chain(18)), ("xxz_ring18", 18, 8, ring(18)),
                                             ("xxz_ladder2x10", 20, 10, square(2, 10)), ("xxz_chain22_n3", 22, 3, chain(22))],
                         ids=["chain18", "ring18_n8", "ladder2x10", "chain22_n3"])
@pytest.mark.parametrize("eng", [0])
def test_long_row_spin_flavour(emu, name, N, n, bonds, eng):
    """The Heisenberg fast path (BASELINE config C3 is the 32-site chain): spin strings of more than
    16 sites through the long-row tables with the spin diagonal."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x = rng.standard_normal(len(orc.enumerate_states(N, n)))
    dz = 0.5
    sd = (dz * len(bonds), -2.0 * dz)
    emu.emu_long_set_spin.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
    emu.emu_long_set_spin(1, sd[0], sd[1])
    try:
        rc, y, cov = run_long(emu, N, n, bonds, 0, 0.0, 0.25, 0, 0.0, eng, x)
    finally:
        emu.emu_long_set_spin(0, 0.0, 0.0)
    if rc == 1:
        pytest.skip("no sub-row class of this sector fits the class-major kernel")
    assert rc == 0, rc
    assert cov.any()
    ref = direct_row(N, n, bonds, 0, 0.0, 0.25, 0, 0.0, x, spin=True, sd=sd)
    assert np.abs(y[cov] - ref[cov]).max() <= 1e-13 * max(1.0, np.abs(ref).max())


def direct_row_general(L, n_dn, bonds, width, hop, u, eps, ups, eu, x):
    """(D + T_dn) x for one row with per-bond hop amplitudes, per-site U and on-site energies."""
    dn = np.asarray(orc.enumerate_states(L, n_dn), dtype=np.int64)
    rank = {int(s): i for i, s in enumerate(dn)}
    y = np.zeros_like(x)
    for i, s in enumerate(dn):
        s = int(s)
        diag = eu + sum(eps[k] for k in range(L) if (s >> k) & 1) + sum(u[k] for k in range(L) if ((s & ups) >> k) & 1)
        acc = 0.0
        for b, (a, c) in enumerate(bonds):
            if ((s >> a) & 1) == ((s >> c) & 1):
                continue
            between = sum(1 << k for k in range(a + 1, c) if k < width)
            sign = -1.0 if bin(s & between).count("1") & 1 else 1.0
            acc += sign * hop[b] * x[rank[s ^ (1 << a) ^ (1 << c)]]
        y[i] = diag * x[i] + acc
    return y


SEG_CASES = [
    ("chain8", 8, 4, chain(8)), ("ring10", 10, 5, ring(10)), ("sq4x3_n5", 12, 5, square(4, 3)),
    ("chain12_n3", 12, 3, chain(12)), ("star12", 12, 6, [(0, j) for j in range(1, 12)]),
    ("chain16", 16, 8, chain(16)), ("sq4x4", 16, 8, square(4, 4)), ("sq4x4_periodic", 16, 7, square(4, 4, True)),
    ("chain5_n2", 5, 2, chain(5)),
]


@pytest.mark.parametrize("name,L,n_dn,bonds", SEG_CASES, ids=[c[0] for c in SEG_CASES])
@pytest.mark.parametrize("uniform", [1, 0])
@pytest.mark.parametrize("width_full", [True, False])
def test_segment_kernel_dn_part(emu, name, L, n_dn, bonds, uniform, width_full):
    """The per-amplitude device function of the default H.v kernel (hub_seg_kernel) on the CPU:
    two-level (dh, dl) tables of build_seg_tables, uniform and site / bond dependent parameters,
    with the fermion sign (width = L) and without (width = 0, Anderson convention)."""
    rng = np.random.default_rng(zlib.crc32((name + str(uniform)).encode()))
    num = len(orc.enumerate_states(L, n_dn))
    x = rng.standard_normal(num)
    ups = int(rng.integers(0, 1 << L))
    width = L if width_full else 0
    if uniform:
        hop, u, eps = np.full(len(bonds), 0.8), np.full(L, 3.5), np.full(L, -0.3)
    else:
        hop, u, eps = rng.uniform(0.5, 1.5, len(bonds)), rng.uniform(0.0, 4.0, L), rng.uniform(-1.0, 1.0, L)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    emu.emu_seg_row.restype = ctypes.c_int
    emu.emu_seg_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, dp, dp, dp,
                                ctypes.c_int, ctypes.c_uint, ctypes.c_double, dp, dp]
    s1 = np.ascontiguousarray([b[0] for b in bonds], dtype=np.int32)
    s2 = np.ascontiguousarray([b[1] for b in bonds], dtype=np.int32)
    hop, u, eps = (np.ascontiguousarray(a, dtype=np.float64) for a in (hop, u, eps))
    y = np.empty_like(x)
    eu = 0.37
    rc = emu.emu_seg_row(L, n_dn, len(bonds), s1.ctypes.data_as(ip), s2.ctypes.data_as(ip), width,
                         hop.ctypes.data_as(dp), u.ctypes.data_as(dp), eps.ctypes.data_as(dp), uniform, ups, eu,
                         x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    if rc == 1:
        pytest.skip("sector outside the segment kernel")
    assert rc == 0, rc
    ref = direct_row_general(L, n_dn, bonds, width, hop, u, eps, ups, eu, x)
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("L,n", [(4, 2), (8, 4), (10, 0), (10, 10), (12, 6), (16, 8), (20, 3), (31, 2), (40, 2), (62, 1), (63, 2)])
def test_sector_unrank_rank(emu, L, n):
    """K1: combinadic unrank / rank (colex order = ascending integers of fixed popcount,
    cmpy/basis.py:655-666 and the bisect of cmpy/operators.py:276-299) on the CPU."""
    llp = ctypes.POINTER(ctypes.c_longlong)
    emu.emu_sector_enumerate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, llp]
    emu.emu_sector_rank.argtypes = [llp, ctypes.c_longlong, llp]
    ref = np.asarray(orc.enumerate_states(L, n), dtype=np.int64)
    count = len(ref)
    out = np.empty(count, dtype=np.int64)
    assert emu.emu_sector_enumerate(L, n, 0, count, out.ctypes.data_as(llp)) == 0
    np.testing.assert_array_equal(out.astype(np.uint64), ref.astype(np.uint64))
    idx = np.empty(count, dtype=np.int64)
    assert emu.emu_sector_rank(out.ctypes.data_as(llp), count, idx.ctypes.data_as(llp)) == 0
    np.testing.assert_array_equal(idx, np.arange(count))


def test_sector_unrank_window_of_the_c5_sector(emu):
    """A window in the middle of C(20, 10) and the last strings of C(32, 16) (config C3)."""
    llp = ctypes.POINTER(ctypes.c_longlong)
    emu.emu_sector_enumerate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, llp]
    emu.emu_sector_rank.argtypes = [llp, ctypes.c_longlong, llp]
    for L, n, first in [(20, 10, 92378 - 50), (32, 16, 601080390 - 100)]:
        out = np.empty(100, dtype=np.int64)
        assert emu.emu_sector_enumerate(L, n, first, 100, out.ctypes.data_as(llp)) == 0
        assert all(bin(int(v)).count("1") == n for v in out) and np.all(np.diff(out) > 0)
        idx = np.empty(100, dtype=np.int64)
        emu.emu_sector_rank(out.ctypes.data_as(llp), 100, idx.ctypes.data_as(llp))
        np.testing.assert_array_equal(idx, first + np.arange(100))
    assert int(out[-1]) == ((1 << 16) - 1) << 16      # the last Sz = 0 string of 32 sites


def test_weighted_elements_bit_exact(emu):
    """K3: sum over set bits in ascending site order (cmpy/operators.py:226-250); plain adds, so the
    result is bit-identical to the Python loop."""
    llp, dp = ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double)
    emu.emu_weighted_elements.argtypes = [llp, ctypes.c_longlong, ctypes.c_int, dp, dp]
    rng = np.random.default_rng(8)
    vals = rng.standard_normal(12)
    states = np.asarray(orc.enumerate_states(12, 5), dtype=np.int64)
    out = np.empty(len(states))
    assert emu.emu_weighted_elements(states.ctypes.data_as(llp), len(states), 12, vals.ctypes.data_as(dp),
                                     out.ctypes.data_as(dp)) == 0
    ref = np.empty(len(states))
    for i, s in enumerate(states):
        v = 0.0
        for k in range(12):
            if (int(s) >> k) & 1:
                v += vals[k]
        ref[i] = v
    np.testing.assert_array_equal(out, ref)


# ---------------------------------------------------------------------------------------
# randomised lattices (hypothesis): arbitrary bond graphs, fillings and sign widths
# ---------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@st.composite
def random_lattice(draw, max_sites=13):
    L = draw(st.integers(4, max_sites))
    n = draw(st.integers(1, L - 1))
    pairs = [(a, b) for a in range(L) for b in range(a + 1, L)]
    bonds = draw(st.lists(st.sampled_from(pairs), min_size=1, max_size=min(len(pairs), 14), unique=True))
    width = draw(st.sampled_from([0, L]))
    ups = draw(st.integers(0, (1 << L) - 1))
    seed = draw(st.integers(0, 2 ** 16))
    return L, n, sorted(bonds), width, ups, seed


@settings(max_examples=40, deadline=None)
@given(random_lattice())
def test_random_lattices_class_major(emu, case):
    L, n, bonds, width, ups, seed = case
    x = np.random.default_rng(seed).standard_normal(len(orc.enumerate_states(L, n)))
    ref = direct_row(L, n, bonds, width, 2.5, -0.9, ups, 0.7, x)
    for eng in (0,):
        rc, y, _ = run_emu(emu, L, n, bonds, width, 2.5, -0.9, ups, 0.7, eng, x)
        if rc == 1:
            continue   # odd row length, or more LH bonds than the class-major tables hold
        assert rc == 0, (rc, case, eng)
        assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (case, eng)


@settings(max_examples=40, deadline=None)
@given(random_lattice(), st.booleans())
def test_random_lattices_segment_kernel(emu, case, uniform):
    L, n, bonds, width, ups, seed = case
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(len(orc.enumerate_states(L, n)))
    if uniform:
        hop, u, eps = np.full(len(bonds), 1.1), np.full(L, 2.0), np.full(L, 0.4)
    else:
        hop, u, eps = rng.uniform(-1.5, 1.5, len(bonds)), rng.uniform(0.0, 4.0, L), rng.uniform(-1.0, 1.0, L)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    emu.emu_seg_row.restype = ctypes.c_int
    emu.emu_seg_row.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, ctypes.c_int, dp, dp, dp,
                                ctypes.c_int, ctypes.c_uint, ctypes.c_double, dp, dp]
    s1 = np.ascontiguousarray([b[0] for b in bonds], dtype=np.int32)
    s2 = np.ascontiguousarray([b[1] for b in bonds], dtype=np.int32)
    hop, u, eps = (np.ascontiguousarray(a, dtype=np.float64) for a in (hop, u, eps))
    y = np.empty_like(x)
    rc = emu.emu_seg_row(L, n, len(bonds), s1.ctypes.data_as(ip), s2.ctypes.data_as(ip), width,
                         hop.ctypes.data_as(dp), u.ctypes.data_as(dp), eps.ctypes.data_as(dp), int(uniform), ups, -0.2,
                         x.ctypes.data_as(dp), y.ctypes.data_as(dp))
    if rc == 1:
        return
    assert rc == 0, (rc, case)
    ref = direct_row_general(L, n, bonds, width, hop, u, eps, ups, -0.2, x)
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), case


@st.composite
def random_long_lattice(draw):
    L = draw(st.integers(17, 20))
    n = draw(st.integers(1, 4))
    pairs = [(a, b) for a in range(L) for b in range(a + 1, L)]
    bonds = draw(st.lists(st.sampled_from(pairs), min_size=1, max_size=12, unique=True))
    width = draw(st.sampled_from([0, L]))
    ups = draw(st.integers(0, (1 << L) - 1))
    seed = draw(st.integers(0, 2 ** 16))
    return L, n, sorted(bonds), width, ups, seed


@settings(max_examples=30, deadline=None)
@given(random_long_lattice())
def test_random_lattices_long_rows(emu, case):
    """Random bond graphs on 17-20 sites: bonds inside the low 16 bits, inside the top bits and
    straddling site 15 | 16 in every mixture."""
    L, n, bonds, width, ups, seed = case
    x = np.random.default_rng(seed).standard_normal(len(orc.enumerate_states(L, n)))
    ref = direct_row(L, n, bonds, width, 3.0, 0.8, ups, -0.4, x)
    for eng in (0,):
        rc, y, cov = run_long(emu, L, n, bonds, width, 3.0, 0.8, ups, -0.4, eng, x)
        if rc == 1:
            continue
        assert rc == 0, (rc, case, eng)
        if cov.any():
            assert np.abs(y[cov] - ref[cov]).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (case, eng)
