"""CPU-only tests: host logic of the package, the C-ABI library loads and exports every
symbol declared in include/cmpy_b200.h, host-only helpers of the library (binomial,
tridiagonal solver), and the no-CPU-fallback contract."""
import ctypes
import os
import re

import numpy as np
import pytest

import cmpy_b200 as cm
from cmpy_b200 import _lib
from conftest import ROOT, has_cuda


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "cmpy_b200.h")).read()
    names = set(re.findall(r"\b(cmpy_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(names):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    assert _lib.lib().cmpy_version() >= 100


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 (plain pointers and sizes only)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                          os.path.join(ROOT, "include", "cmpy_b200.h")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_c_example_compiles_and_links():
    """examples/e0_from_c.c: a plain-C client of the ABI (no Python, no PyTorch) builds against the
    header and links against the shared library."""
    import shutil
    import subprocess
    import tempfile

    cuda_inc, cuda_lib = "/usr/local/cuda/include", "/usr/local/cuda/lib64"
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    with tempfile.TemporaryDirectory() as tmp:
        res = subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                              os.path.join(ROOT, "examples", "e0_from_c.c"), "-L", os.path.dirname(_lib.LIB_PATH),
                              "-lcmpy_b200", "-L", cuda_lib, "-lcudart", "-lm", "-o", os.path.join(tmp, "e0")],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_binomial_and_errors():
    assert _lib.binomial(20, 10) == 184756
    assert _lib.binomial(32, 16) == 601080390
    assert _lib.binomial(5, 7) == 0
    with pytest.raises(ValueError):
        _lib.binomial(100, 3)
    assert "binomial" in _lib.last_error()


def test_tridiag_solver_host():
    from cmpy_b200.exactdiag import lanczos_ground_state, lanczos_matrix

    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 200):
        a = rng.standard_normal(n)
        b = rng.uniform(0.1, 1.0, size=max(n - 1, 0))
        e, v = lanczos_ground_state(a, b)
        t = lanczos_matrix(a, b) if n > 1 else np.array([[a[0]]])
        w = np.linalg.eigvalsh(t)
        assert abs(e - w[0]) < 1e-12 * max(1.0, abs(w).max())
        assert np.abs(t @ v - e * v).max() < 1e-9
        assert abs(np.linalg.norm(v) - 1.0) < 1e-12


def test_bit_helpers_match_reference_semantics():
    # cmpy/tests/test_basis.py:14-54 (f-string semantics)
    for num in (0, 1, 5, 6, 37, 255):
        for width in (0, 3, 9):
            s = f"{num:0{width}b}"
            assert cm.binstr(num, width) == s
            assert list(cm.binarr(num, width)) == [int(c) for c in s[::-1]]
            assert cm.binidx(num, width) == [i for i, c in enumerate(s[::-1]) if c == "1"]
            assert list(cm.overlap(num, 0b101, width)) == list(cm.binarr(num & 0b101, width))
    assert cm.create(0b0101, 1) == 0b0111 and cm.create(0b0101, 0) is None
    assert cm.annihilate(0b0101, 0) == 0b0100 and cm.annihilate(0b0101, 1) is None
    from cmpy_b200.basis import upper_sector, lower_sector, get_ibit, set_ibit

    assert upper_sector(2, 3, cm.UP, 4) == (3, 3) and upper_sector(4, 3, cm.UP, 4) is None
    assert upper_sector(2, 4, cm.DN, 4) is None and lower_sector(0, 1, cm.UP) is None
    assert lower_sector(2, 1, cm.DN) == (2, 0)
    assert get_ibit(0b1101, 2, 2) == 3 and set_ibit(0b0001, 1, 1) == 3
    assert cm.state_label(0b01, 0b11, 3) == "⇅↓."


def test_basis_host_parts():
    b = cm.Basis(5, init_sectors=True)
    assert b.num_spinstates == 32 and b.size == 1024 and b.fillings == [0, 1, 2, 3, 4, 5]
    assert all(x.bit_count() == 2 for x in b.sectors[2]) and len(b.sectors[2]) == 10
    b = cm.Basis(4)
    assert b.get_states(None) == list(range(16)) and b.get_states(0) == [0]
    assert b.get_states(1) == [1, 2, 4, 8]
    sec = b.get_sector(1, None)
    assert sec.size == 64 and sec.num_dn == 16
    assert [s.up for s in list(sec.states)[:2]] == [1, 1]
    assert b.upper_sector(4, 0, cm.UP) is None
    sb = cm.SpinBasis(4)
    assert sb.num_up(0) == 2 and sb.num_up(-1) == 1
    with pytest.raises(ValueError):
        cm.SpinBasis(3).num_up(0)
    # index layout (cmpy/tests/test_operator.py:15-28): idx = up_idx * num_dn + dn_idx
    full = cm.Basis(3).get_sector()
    for up_idx in (0, 3, 5):
        idx = [i for i, st in enumerate(full.states) if st.up == up_idx]
        assert list(cm.project_up(up_idx, full.num_dn, np.arange(full.num_dn))) == idx
    for dn_idx in (0, 2, 7):
        idx = [i for i, st in enumerate(full.states) if st.dn == dn_idx]
        assert list(cm.project_dn(dn_idx, full.num_dn, np.arange(full.num_up))) == idx
    assert list(cm.project_elements_up(1, 4, np.arange(4), 1.0, target=2))[0] == (4, 8, 1.0)


def test_model_parameters_and_specs():
    from cmpy_b200.models import HubbardModel, SingleImpurityAndersonModel, HeisenbergModel
    from refshim import ChainStandIn

    m = HubbardModel(4, [[0, 1], [1, 2], [3, 2]], inter=4.0, mu=2.0)
    assert m.inter == 4.0 and m["mu"] == 2.0 and m.num_sites == 4
    m.hop = 0.5
    assert m.params["hop"] == 0.5
    spec = m._operator_spec()
    assert spec["bonds"] == [(0, 1), (1, 2)] and spec["sign_width"] == 4  # i >= j dropped
    assert list(spec["eps"]) == [-2.0] * 4
    siam = SingleImpurityAndersonModel(u=2.0, eps_bath=[0.1, 0.2], v=1.0)
    assert siam.mu == 1.0 and siam.num_sites == 3 and list(siam.v) == [1.0, 1.0]
    assert siam._operator_spec()["sign_width"] == 0
    with pytest.raises(AssertionError):
        SingleImpurityAndersonModel(eps_bath=[0.1, 0.2], v=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        siam.update_hybridization([1.0])
    h = HeisenbergModel(ChainStandIn(4), j=2.0)
    assert h.jz == 2.0 and len(h._operator_spec()["pairs"]) == 6
    assert cm.EigenState().energy == np.inf


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from cmpy_b200.models import HubbardModel

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cm.Basis(6).get_states(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HubbardModel(2, [[0, 1]]).hamilton_operator(1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cm.HamiltonOperator(2, [1.0], [[0], [1]])


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = ("import sys; import cmpy_b200, cmpy_b200.exactdiag, cmpy_b200.greens, cmpy_b200.models;"
            "bad=[m for m in sys.modules if m.startswith('oracle') or m=='refshim'];"
            "assert not bad, bad")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cmpy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_np" not in src and "refshim" not in src, f


def test_operator_spec_follows_the_generator_hook():
    """A subclass that changes `_hamiltonian_data` without describing itself again must not inherit the
    parent's matrix-free description (round-1 advisor finding; reference hook: cmpy/models/abc.py:238-256)."""
    from cmpy_b200.models import HubbardModel

    class Extended(HubbardModel):
        def _hamiltonian_data(self, up_states, dn_states):
            yield from super()._hamiltonian_data(up_states, dn_states)
            yield 0, 0, 1.0

    class Redescribed(Extended):
        def _operator_spec(self):
            return HubbardModel._operator_spec(self)

    nb = [[0, 1], [1, 2]]
    base = HubbardModel(3, nb, inter=2.0, mu=1.0, hop=1.0)
    assert base._trusted_spec() is not None
    assert Extended(3, nb, inter=2.0, mu=1.0, hop=1.0)._trusted_spec() is None
    assert Redescribed(3, nb, inter=2.0, mu=1.0, hop=1.0)._trusted_spec() is not None
    # the parameter container still behaves like the reference's (cmpy/models/abc.py:21-133)
    assert base.inter == 2.0 and base["inter"] == 2.0
    base.inter = 3.0
    assert base["inter"] == 3.0 and "inter" in base.params
    assert str(base).startswith("HubbardModel(") and "U=3.0" in base.pformat()
    from cmpy_b200.models.abc import ModelParameters

    mp = ModelParameters(a=1.0, b=2)
    mp.rename_param("a", "c")
    assert mp.c == 1.0 and "a" not in mp and list(mp) == ["b", "c"]
    assert mp.key(decimals=1) == "b=2.0; c=1.0" and mp.json() == '{"b": 2, "c": 1.0}'
    mp.delete_param("b")
    assert len(mp) == 1 and mp.pformat() == "c=1.0"
    with pytest.raises(AttributeError):
        mp.nope
