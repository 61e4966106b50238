"""Host-side helpers under the property / known-answer tests the reference ships for them
(cmpy/tests/test_basis.py:14-141, cmpy/tests/test_operator.py:15-28), run against cmpy_b200 with
the same strategies and expected values.  No GPU: these functions are pure host logic."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from numpy.testing import assert_array_equal

import cmpy_b200 as cm
from cmpy_b200 import basis

WIDTHS = [0, 5, 10, 15, 20]
NUM = st.integers(0, 2 ** 15)


def bits_lsb_first(num, width=0):
    return np.fromiter(f"{num:0{width}b}"[::-1], dtype=np.int64)


@given(NUM)
def test_binstr(num):
    assert basis.binstr(num) == f"{num:b}"
    for width in WIDTHS:
        assert basis.binstr(num, width) == f"{num:0{width}b}"


@given(NUM)
def test_binarr_and_occupations(num):
    assert_array_equal(basis.binarr(num), bits_lsb_first(num))
    assert_array_equal(basis.occupations(num), bits_lsb_first(num))
    for width in WIDTHS:
        assert_array_equal(basis.binarr(num, width), bits_lsb_first(num, width))
        assert_array_equal(basis.occupations(num, width), bits_lsb_first(num, width))


@given(NUM)
def test_binidx(num):
    assert_array_equal(basis.binidx(num), np.where(basis.binarr(num))[0])
    for width in WIDTHS:
        assert_array_equal(basis.binidx(num, width), np.where(basis.binarr(num, width))[0])


@given(NUM, NUM)
def test_overlap(num1, num2):
    assert_array_equal(basis.overlap(num1, num2), bits_lsb_first(num1 & num2))
    for width in WIDTHS:
        assert_array_equal(basis.overlap(num1, num2, width), bits_lsb_first(num1 & num2, width))


@pytest.mark.parametrize("num,pos,result", [
    (0b0, 0, 0b1), (0b1, 0, None), (0b100, 0, 0b101), (0b100, 1, 0b110),
    (0b100, 2, None), (0b110, 0, 0b111), (0b110, 1, None), (0b110, 2, None),
])
def test_create(num, pos, result):
    assert basis.create(num, pos) == result


@pytest.mark.parametrize("num,pos,result", [
    (0b0, 0, None), (0b1, 0, 0b0), (0b100, 0, None), (0b100, 1, None),
    (0b100, 2, 0b000), (0b110, 0, None), (0b110, 1, 0b100), (0b110, 2, 0b010),
])
def test_annihilate(num, pos, result):
    assert basis.annihilate(num, pos) == result


@pytest.mark.parametrize("n_up", range(15))
@pytest.mark.parametrize("n_dn", range(15))
def test_upper_and_lower_sector(n_up, n_dn):
    num_sites = 15
    assert basis.upper_sector(n_up, n_dn, basis.UP, num_sites) == (None if n_up == num_sites else (n_up + 1, n_dn))
    assert basis.upper_sector(n_up, n_dn, basis.DN, num_sites) == (None if n_dn == num_sites else (n_up, n_dn + 1))
    assert basis.lower_sector(n_up, n_dn, basis.UP) == (None if n_up == 0 else (n_up - 1, n_dn))
    assert basis.lower_sector(n_up, n_dn, basis.DN) == (None if n_dn == 0 else (n_up, n_dn - 1))


@given(st.integers(0, 15))
def test_basis_sizes(num_sites):
    b = basis.Basis(num_sites)
    assert b.num_spinstates == 2 ** num_sites
    assert b.size == 2 ** (2 * num_sites)


@pytest.mark.parametrize("num_sites,n,result", [
    (2, 0, ["00"]), (2, 1, ["01", "10"]), (2, None, ["00", "01", "10", "11"]),
    (3, 0, ["000"]), (3, 1, ["001", "010", "100"]),
    (3, None, ["000", "001", "010", "011", "100", "101", "110", "111"]),
])
def test_basis_get_states_host_cases(num_sites, n, result):
    """The n in (None, 0, 1) branches are host lists (cmpy/basis.py:658-662); the n >= 2 rows of the
    reference table need the enumeration kernel and live in tests/test_gpu_parity.py."""
    states = basis.Basis(num_sites).get_states(n)
    assert isinstance(states, list)
    assert [basis.binstr(s, num_sites) for s in states] == result


@settings(max_examples=12, deadline=None)
@given(st.integers(0, 5))
def test_project_up(up_idx):
    sec = cm.Basis(5).get_sector()
    indices = [i for i, state in enumerate(sec.states) if state.up == up_idx]
    assert_array_equal(indices, cm.project_up(up_idx, sec.num_dn, np.arange(sec.num_dn)))


@settings(max_examples=12, deadline=None)
@given(st.integers(0, 5))
def test_project_dn(dn_idx):
    sec = cm.Basis(5).get_sector()
    indices = [i for i, state in enumerate(sec.states) if state.dn == dn_idx]
    assert_array_equal(indices, cm.project_dn(dn_idx, sec.num_dn, np.arange(sec.num_up)))
