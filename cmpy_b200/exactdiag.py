# -*- coding: utf-8 -*-
"""Exact-diagonalisation entry points with cmpy's interface (reference: cmpy/exactdiag.py).

What is different underneath:

* ``compute_groundstate`` uses the GPU Lanczos driver (``cmpy_lanczos_run``: fused H.v +
  vector kernels, CUDA-graph replay) where the reference calls ARPACK
  ``eigsh(hamop, k=1, which="SA")`` (cmpy/exactdiag.py:37).
* ``gf_continued_fraction`` is the zero-temperature G(z): ground state by Lanczos, a second
  Lanczos run from ``c^dagger|gs>`` / ``c|gs>`` and a continued-fraction evaluation kernel.
  The reference has no such routine; its results are checked against the Lehmann sum built
  from the reference's parts.
* ``gf_lehmann`` keeps the reference's finite-temperature all-sector Lehmann sum
  (cmpy/exactdiag.py:215-245); the dense eigensolves run through cuSOLVER (torch), the
  ``c^dagger`` application and the pole sum are custom kernels.
"""
import ctypes
import logging
from ctypes import POINTER, c_double, c_int

import numpy as np

from . import _lib
from .basis import Sector, UP
from .matrix import EigenState
from .operators import CreationOperator, AnnihilationOperator, HamiltonOperator

logger = logging.getLogger(__name__)

__all__ = [
    "compute_groundstate", "solve_sector", "GreensFunctionMeasurement", "gf_lehmann",
    "gf_continued_fraction", "lanczos_run", "lanczos_groundstate", "LanczosResult",
    "iter_lanczos_coeffs", "lanczos_coeffs", "lanczos_matrix", "lanczos_ground_state",
    "cf_eval", "pole_sum", "gf_greater", "gf_lesser", "gf_tevo", "fourier_t2z",
]


# =========================================================================================
# GPU Lanczos
# =========================================================================================

class LanczosResult:
    """alpha[m], beta[m+1] (beta[0] = |v0|, beta[k] = norm of the k-th residual), lowest
    Ritz value, residual estimate, and optionally the normalised Ritz vector (CUDA tensor)."""

    __slots__ = ["alpha", "beta", "nit", "e0", "resid", "vector", "converged"]

    def __init__(self, alpha, beta, nit, e0, resid, vector, converged):
        self.alpha, self.beta, self.nit, self.e0 = alpha, beta, nit, e0
        self.resid, self.vector, self.converged = resid, vector, converged


def _start_vector(size, seed=0):
    """Seeded standard-normal start vector, normalised, generated on the device."""
    torch = _lib.require_cuda()
    g = torch.Generator(device=_lib.device())
    g.manual_seed(seed)
    v = torch.randn(size, dtype=torch.float64, device=_lib.device(), generator=g)
    return v / v.norm()


def lanczos_run(hamop, v0=None, maxit=500, tol=1e-10, resid_tol=0.0, check_every=10,
                want_vector=False, use_graph=True, seed=0):
    """Plain Lanczos on a device operator (``HamiltonOperator`` family).

    ``v0``: CUDA float64 tensor or numpy vector (default: seeded standard-normal vector).
    Stops when the lowest Ritz value moves by less than ``tol`` between checks *and* the
    residual estimate is below ``resid_tol`` (``<= 0``: residual^2/gap < tol), or after
    ``maxit`` iterations (then ``converged`` is False)."""
    torch = _lib.require_cuda()
    n = hamop.shape[0]
    if v0 is None:
        v0 = _start_vector(n, seed)
    elif not isinstance(v0, torch.Tensor):
        v0 = torch.from_numpy(np.ascontiguousarray(v0, dtype=np.float64)).to(_lib.device())
    v0 = v0.contiguous()
    dev = v0.device
    w0 = torch.empty(n, dtype=torch.float64, device=dev)
    w1 = torch.empty(n, dtype=torch.float64, device=dev)
    vec = torch.empty(n, dtype=torch.float64, device=dev) if want_vector else None
    cap = int(maxit) + int(check_every) + 4
    alpha = np.zeros(cap, dtype=np.float64)
    beta = np.zeros(cap + 1, dtype=np.float64)
    nit, e0, resid = c_int(0), c_double(0.0), c_double(0.0)
    rc = _lib.lib().cmpy_lanczos_run(
        hamop.handle, _lib.ptr(v0), _lib.ptr(w0), _lib.ptr(w1), int(maxit), float(tol),
        float(resid_tol), int(check_every), int(bool(use_graph)),
        alpha.ctypes.data_as(POINTER(c_double)), beta.ctypes.data_as(POINTER(c_double)),
        ctypes.byref(nit), ctypes.byref(e0), ctypes.byref(resid), _lib.ptr(vec), _lib.stream_ptr())
    converged = rc == _lib.CMPY_OK
    if rc not in (_lib.CMPY_OK, _lib.CMPY_ERR_NOT_CONVERGED):
        _lib.check(rc, "cmpy_lanczos_run")
    m = nit.value
    return LanczosResult(alpha[:m].copy(), beta[:m + 1].copy(), m, e0.value, resid.value, vec,
                         converged)


def lanczos_groundstate(hamop, tol=1e-10, resid_tol=1e-9, maxit=2000, want_vector=True, seed=0):
    """Lowest eigenpair of a device operator: ``(energy, state)`` with ``state`` a CUDA
    tensor (or ``None``). Replaces ``sla.eigsh(hamop, k=1, which="SA")`` (exactdiag.py:37)."""
    res = lanczos_run(hamop, None, maxit=maxit, tol=tol, resid_tol=resid_tol, check_every=10,
                      want_vector=want_vector, seed=seed)
    if not res.converged:
        logger.warning("Lanczos did not converge in %d iterations (resid %.2e)", res.nit, res.resid)
    return res.e0, res.vector


def compute_groundstate(model, thresh=50):
    """Lowest eigenstate over all (n_up, n_dn) sectors (reference: cmpy/exactdiag.py:24-43).
    Sectors up to ``thresh`` states are diagonalised densely as in the reference; larger ones
    by GPU Lanczos.  ``state`` is returned as a numpy vector."""
    gs = EigenState()
    logger.debug("Computing ground-state:")
    for n_up, n_dn in model.basis.iter_fillings():
        hamop = model.hamilton_operator(n_up=n_up, n_dn=n_dn, dtype=np.float64)
        logger.debug("Sector (%d, %d), size: %d", n_up, n_dn, hamop.shape[0])
        if hamop.shape[0] <= thresh:
            energies, vectors = np.linalg.eigh(hamop.toarray())
            idx = np.argmin(energies)
            energy, state = energies[idx], vectors[:, idx]
        else:
            energy, vec = lanczos_groundstate(hamop)
            state = vec.cpu().numpy()
        if energy < gs.energy:
            gs = EigenState(energy, state, n_up, n_dn)
            logger.debug("gs-energy: %.6f", gs.energy)
    return gs


# =========================================================================================
# continued fraction / pole sums (K8)
# =========================================================================================

def _z_tensor(z):
    torch = _lib.require_cuda()
    zz = np.ascontiguousarray(np.atleast_1d(z), dtype=np.complex128)
    return torch.from_numpy(zz).to(_lib.device()), zz.shape


def cf_eval(alpha, beta, norm2, e0, z, sign=+1, out=None):
    """``norm2 / (z - s(a0-e0) - b1^2/(z - s(a1-e0) - ...))`` on the device (kernel K8).
    ``beta`` holds the off-diagonals b1..b_{m-1}. Returns a CUDA complex128 tensor."""
    torch = _lib.require_cuda()
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    b = np.ascontiguousarray(beta, dtype=np.float64)
    if len(a) < 1 or len(b) < len(a) - 1:
        raise ValueError("need len(beta) >= len(alpha) - 1 >= 0")
    zt = z if isinstance(z, torch.Tensor) else _z_tensor(z)[0]
    accumulate = out is not None
    if out is None:
        out = torch.empty_like(zt)
    _lib.check(_lib.lib().cmpy_cf_eval(
        a.ctypes.data_as(POINTER(c_double)), b.ctypes.data_as(POINTER(c_double)), len(a),
        float(norm2), float(e0), int(sign), _lib.ptr(zt), zt.numel(), _lib.ptr(out),
        int(accumulate), _lib.stream_ptr()), "cmpy_cf_eval")
    return out


def pole_sum(weights, poles, z, out=None):
    """``sum_k w_k / (z - p_k)`` on the device; weights/poles: CUDA float64 tensors."""
    torch = _lib.require_cuda()
    zt = z if isinstance(z, torch.Tensor) else _z_tensor(z)[0]
    accumulate = out is not None
    if out is None:
        out = torch.empty_like(zt)
    w = weights.contiguous().view(-1)
    p = poles.contiguous().view(-1)
    _lib.check(_lib.lib().cmpy_pole_sum(_lib.ptr(w), _lib.ptr(p), w.numel(), _lib.ptr(zt),
                                        zt.numel(), _lib.ptr(out), int(accumulate),
                                        _lib.stream_ptr()), "cmpy_pole_sum")
    return out


def gf_continued_fraction(model, z, pos=0, sigma=UP, n_up=None, n_dn=None, gs=None, num_coeffs=600,
                          tol=1e-12, signed=False, return_info=False):
    """Zero-temperature G_{pos,sigma}(z) = <gs|c 1/(z-(H-E0)) c^+|gs> + <gs|c^+ 1/(z+(H-E0)) c|gs>
    by Lanczos + continued fraction, entirely on the device.

    ``pos`` is a site or a pair ``(i, j)`` (off-diagonal G_ij, use ``signed=True`` there).
    The ground state is searched in sector ``(n_up, n_dn)`` (default: half filling
    ``num_sites//2`` each).  ``signed=False`` uses the reference's signless ladder operators
    (cmpy/operators.py:652-703), so the result matches ``gf_lehmann`` in the T -> 0 limit."""
    torch = _lib.require_cuda()
    basis = model.basis
    L = basis.num_sites
    n_up = L // 2 if n_up is None else n_up
    n_dn = L // 2 if n_dn is None else n_dn
    sector = basis.get_sector(n_up, n_dn)
    if gs is None:
        hamop = model.hamilton_operator(sector=sector)
        res = lanczos_run(hamop, None, maxit=3000, tol=tol, resid_tol=1e-10, check_every=10,
                          want_vector=True)
        e0, psi = res.e0, res.vector
    else:
        e0, psi = gs
        if not isinstance(psi, torch.Tensor):
            psi = torch.from_numpy(np.ascontiguousarray(psi, dtype=np.float64)).to(_lib.device())
    zt, zshape = _z_tensor(z)
    g = torch.zeros_like(zt)
    info = {"e0": e0, "norms": [0.0, 0.0], "nit": [0, 0]}
    pair = isinstance(pos, (tuple, list))
    if pair and pos[0] == pos[1]:
        pos, pair = int(pos[0]), False
    for part, (sec_t, sign) in enumerate(((basis.upper_sector(n_up, n_dn, sigma), +1),
                                           (basis.lower_sector(n_up, n_dn, sigma), -1))):
        if sec_t is None:
            continue
        cls = CreationOperator if sign > 0 else AnnihilationOperator
        ham_t = None
        if pair:
            # off-diagonal G_ij (reference sketch: lehmann_full.py:53-97) by
            # polarisation: for real symmetric H and a real ground state
            #   <T_i gs| R(z) |T_j gs> = ( <phi+|R|phi+> - <phi-|R|phi-> ) / 4,   phi+- = (T_i +- T_j)|gs>
            pi_ = cls(sector, sec_t, pos=int(pos[0]), sigma=sigma, signed=signed).apply(psi)
            pj_ = cls(sector, sec_t, pos=int(pos[1]), sigma=sigma, signed=signed).apply(psi)
            starts = ((pi_ + pj_, 0.25), (pi_ - pj_, -0.25))
        else:
            starts = ((cls(sector, sec_t, pos=pos, sigma=sigma, signed=signed).apply(psi), 1.0),)
        for phi, weight in starts:
            norm2 = float(torch.dot(phi, phi))
            info["norms"][part] += weight * norm2
            if norm2 < 1e-28:
                continue
            if ham_t is None:
                ham_t = model.hamilton_operator(sector=sec_t)
            m = min(num_coeffs, ham_t.shape[0])
            res = lanczos_run(ham_t, phi, maxit=m, tol=0.0, resid_tol=0.0, check_every=50)
            info["nit"][part] = max(info["nit"][part], res.nit)
            cf_eval(res.alpha, res.beta[1:res.nit], weight * norm2, e0, zt, sign=sign, out=g)
    out = g.cpu().numpy().reshape(zshape)
    return (out, info) if return_info else out


# =========================================================================================
# finite-temperature Lehmann sum (reference: cmpy/exactdiag.py:46-245)
# =========================================================================================

def solve_sector(model, sector: Sector, cache: dict = None):
    """Full spectrum of one sector: dense Hamiltonian from the device operator, cuSOLVER
    ``eigh`` (the reference uses LAPACK, cmpy/exactdiag.py:46-57). Returns numpy arrays."""
    torch = _lib.require_cuda()
    key = (sector.n_up, sector.n_dn)
    if cache is not None and key in cache:
        return cache[key]
    logger.debug("Solving eig  %s, %s (%s)", sector.n_up, sector.n_dn, sector.size)
    ham = torch.from_numpy(model.hamiltonian(sector=sector)).to(_lib.device())
    eigvals, eigvecs = torch.linalg.eigh(ham)
    result = [eigvals.cpu().numpy(), eigvecs.cpu().numpy()]
    if cache is not None:
        cache[key] = result
    return result


class GreensFunctionMeasurement:
    """Accumulator of the Lehmann sum, partition function and occupations with the running
    ground-state rescaling of the reference (cmpy/exactdiag.py:137-212)."""

    def __init__(self, z, beta, pos=0, sigma=UP, dtype=None, measure_occ=True):
        self.z = z
        self.beta = beta
        self.pos = pos
        self.sigma = sigma
        self._measure_occ = measure_occ
        self._part = 0
        self._gs_energy = np.inf
        self._gf = np.zeros_like(z, dtype=dtype)
        self._occ = 0.0
        self._occ_double = 0.0

    @property
    def part(self):
        return self._part * np.exp(-self.beta * self._gs_energy)

    @property
    def gf(self):
        return self._gf / self._part

    @property
    def occ(self):
        return self._occ / self._part

    @property
    def occ_double(self):
        return self._occ_double / self._part

    @property
    def gs_energy(self):
        return self._gs_energy

    def accumulate(self, sector, sector_p1, evals, evecs, evals_p1, evecs_p1):
        torch = _lib.require_cuda()
        dev = _lib.device()
        beta = self.beta
        min_energy = min(evals)
        factor = 1.0
        if min_energy < self._gs_energy:
            factor = np.exp(-beta * (self._gs_energy - min_energy))
            self._gs_energy = min_energy
        e0 = self._gs_energy
        self._part = self._part * factor + np.sum(np.exp(-beta * (evals - e0)))
        if factor != 1.0:
            self._gf *= factor
        # overlap[m, n] = |<m| c^dagger |n>|^2 : ladder kernel on every eigenvector + GEMM
        cdag = CreationOperator(sector, sector_p1, pos=self.pos, sigma=self.sigma)
        ev = torch.from_numpy(np.ascontiguousarray(evecs.T)).to(dev)  # rows = eigenvectors
        cd_ev = torch.empty((ev.shape[0], cdag.shape[0]), dtype=torch.float64, device=dev)
        for k in range(ev.shape[0]):
            cdag.apply(ev[k], out=cd_ev[k])
        ev1 = torch.from_numpy(np.ascontiguousarray(evecs_p1)).to(dev)
        overlap = (cd_ev @ ev1).T.abs() ** 2  # (m, n)
        ex = torch.from_numpy(np.exp(-beta * (evals - e0))).to(dev)
        ex1 = torch.from_numpy(np.exp(-beta * (evals_p1 - e0))).to(dev)
        weights = overlap * (ex[None, :] + ex1[:, None])
        e_n = torch.from_numpy(np.ascontiguousarray(evals)).to(dev)
        e_m = torch.from_numpy(np.ascontiguousarray(evals_p1)).to(dev)
        poles = e_m[:, None] - e_n[None, :]
        zt, zshape = _z_tensor(self.z)
        self._gf = self._gf + pole_sum(weights, poles, zt).cpu().numpy().reshape(zshape)
        if self._measure_occ:
            up = np.asarray(sector.up_states, dtype=np.int64)
            dn = np.asarray(sector.dn_states, dtype=np.int64)
            w_state = ((ev ** 2) * ex[:, None]).sum(dim=0).cpu().numpy().reshape(len(up), len(dn))
            if self.sigma == UP:
                occ = w_state[((up >> self.pos) & 1).astype(bool), :].sum()
            else:
                occ = w_state[:, ((dn >> self.pos) & 1).astype(bool)].sum()
            both = ((up[:, None] & dn[None, :]) >> self.pos) & 1
            self._occ = self._occ * factor + occ
            self._occ_double = self._occ_double * factor + (w_state * both).sum()


def gf_lehmann(model, z, beta, pos=0, sigma=UP, eig_cache=None, occ=True):
    """Finite-temperature Lehmann Green's function over all sector pairs
    (reference: cmpy/exactdiag.py:215-245)."""
    basis = model.basis
    data = GreensFunctionMeasurement(z, beta, pos, sigma, measure_occ=occ)
    eig_cache = eig_cache if eig_cache is not None else dict()
    for n_up, n_dn in basis.iter_fillings():
        sector = model.get_sector(n_up, n_dn)
        sector_p1 = basis.upper_sector(n_up, n_dn, sigma)
        if sector_p1 is not None:
            eigvals, eigvecs = solve_sector(model, sector, cache=eig_cache)
            eigvals_p1, eigvecs_p1 = solve_sector(model, sector_p1, cache=eig_cache)
            data.accumulate(sector, sector_p1, eigvals, eigvecs, eigvals_p1, eigvecs_p1)
    logger.info("gs-energy:  %+.4f", data.gs_energy)
    return data


# =========================================================================================
# the reference's hand-rolled Lanczos interface (cmpy/exactdiag.py:324-375)
# =========================================================================================

def _as_device_operator(ham):
    if hasattr(ham, "handle"):
        return ham
    dense = np.asarray(ham, dtype=np.float64)
    rows, cols = np.nonzero(dense)
    # HamiltonOperator computes y[col] += val * x[row]: pass (col, row) so that y = ham @ x
    return HamiltonOperator(dense.shape[0], dense[rows, cols], np.array([cols, rows]))


def iter_lanczos_coeffs(ham, size=10):
    """Yields ``(a_n, b_n)`` (``b_0`` is ``None``) of the Lanczos recurrence started from
    ``np.random.uniform(0, 1)`` (global RNG, as the reference).  ``ham`` may be a dense array
    (as in the reference) or a device operator; the recurrence runs on the GPU in the
    normalised basis, which yields the same coefficients."""
    op = _as_device_operator(ham)
    psi = np.random.uniform(0, 1, size=op.shape[0])
    res = lanczos_run(op, psi, maxit=size, tol=0.0, resid_tol=0.0, check_every=max(2, size))
    for n in range(min(size, res.nit)):
        yield res.alpha[n], (None if n == 0 else res.beta[n])


def lanczos_coeffs(ham, size=10):
    a_coeffs, b_coeffs = list(), list()
    for a, b in iter_lanczos_coeffs(ham, size):
        a_coeffs.append(a)
        b_coeffs.append(b)
    b_coeffs.pop(0)
    return a_coeffs, b_coeffs


def lanczos_matrix(a_coeffs, b_coeffs):
    mat = np.diag(a_coeffs)
    np.fill_diagonal(mat[1:], b_coeffs)
    np.fill_diagonal(mat[:, 1:], b_coeffs)
    return mat


def lanczos_ground_state(a_coeffs, b_coeffs, max_eig=3):
    """Lowest eigenpair of the Lanczos tridiagonal matrix (bisection + inverse iteration in
    ``cmpy_tridiag_lowest``; reference: scipy ``eigh_tridiagonal``, exactdiag.py:368-375)."""
    a = np.ascontiguousarray(a_coeffs, dtype=np.float64)
    b = np.ascontiguousarray(b_coeffs, dtype=np.float64)
    k = min(max_eig + 1, len(a))
    evals = np.zeros(k)
    vec = np.zeros(len(a))
    _lib.check(_lib.lib().cmpy_tridiag_lowest(
        a.ctypes.data_as(POINTER(c_double)), b.ctypes.data_as(POINTER(c_double)), len(a), k,
        evals.ctypes.data_as(POINTER(c_double)), vec.ctypes.data_as(POINTER(c_double))),
        "cmpy_tridiag_lowest")
    return evals[0], vec


# =========================================================================================
# real-time Green's functions (row f-2 of SURVEY.md section 8(f))
# =========================================================================================
# The reference propagates T|gs> with a Taylor-series expm_multiply on the operator
# (cmpy/exactdiag.py:248-308, cmpy/linalg/expm_multiply.py) and takes the overlap with <gs|T.
# Only the autocorrelation <phi| exp(-/+ i H t) |phi> is needed, so here the GPU Lanczos that
# already feeds the continued fraction is reused: with T_m the Lanczos matrix of (H, phi),
# <phi| f(H) |phi> = |phi|^2 (f(T_m))_00 = |phi|^2 sum_k w_k f(theta_k)  (Gauss quadrature of the
# spectral measure); m is raised until the time series stops changing.

def _krylov_measure(ham_t, phi, tmax, tol=1e-10, m0=200):
    """Ritz values theta_k and weights w_k of the spectral measure of ``phi`` under ``ham_t``."""
    from scipy.linalg import eigh_tridiagonal

    dim = ham_t.shape[0]

    def measure(alpha, beta_off):
        if len(alpha) == 1:
            return np.array([alpha[0]]), np.array([1.0])
        theta, vecs = eigh_tridiagonal(alpha, beta_off)
        return theta, vecs[0] ** 2

    m = min(int(m0), dim)
    prev = None
    probe_t = np.linspace(0.0, float(tmax), 33)
    while True:
        res = lanczos_run(ham_t, phi, maxit=m, tol=0.0, resid_tol=0.0, check_every=max(m, 50), use_graph=False)
        nit = res.nit
        alpha, beta_off = np.array(res.alpha[:nit]), np.array(res.beta[1:nit])
        scale = max(1.0, float(np.abs(alpha).max()))
        small = np.nonzero(beta_off < 1e-12 * scale)[0]
        if small.size:  # invariant subspace reached: the measure is exact
            k = int(small[0]) + 1
            return measure(alpha[:k], beta_off[:k - 1])
        theta, w = measure(alpha, beta_off)
        cur = (w[None, :] * np.exp(-1j * np.outer(probe_t, theta))).sum(axis=1)
        if nit >= dim or (prev is not None and np.abs(cur - prev).max() < tol):
            return theta, w
        prev = cur
        m = min(2 * m, dim)


def _autocorrelation(model, gs, sector_t, ladder_cls, sector, start, stop, num, pos, sigma, direction):
    """times, direction*i * exp(-direction*i*E0*t) ... see gf_greater / gf_lesser."""
    torch = _lib.require_cuda()
    times = np.linspace(start, stop, num)
    psi = gs.state
    if not isinstance(psi, torch.Tensor):
        psi = torch.from_numpy(np.ascontiguousarray(np.real(psi), dtype=np.float64)).to(_lib.device())
    phi = ladder_cls(sector, sector_t, pos=pos, sigma=sigma).apply(psi)
    norm2 = float(torch.dot(phi, phi))
    if norm2 < 1e-28:
        return times, np.zeros(num, dtype=np.complex128)
    ham_t = model.hamilton_operator(sector=sector_t)
    theta, w = _krylov_measure(ham_t, phi, max(abs(start), abs(stop)))
    # G(t_n) = (-/+ i) |phi|^2 sum_k w_k exp(-/+ i (theta_k - E0) t_n), evaluated on the device
    dev = _lib.device()
    t_d = torch.from_numpy(times).to(dev)
    om = torch.from_numpy(theta - float(gs.energy)).to(dev)
    w_d = torch.from_numpy(w * norm2).to(dev).to(torch.complex128)
    phase = torch.exp(-1j * direction * torch.outer(t_d, om).to(torch.complex128))
    g = (-1j * direction) * (phase @ w_d)
    return times, g.cpu().numpy()


def gf_greater(model, gs, start, stop, num=1000, pos=0, sigma=UP):
    """G^>(t) = -i <gs| c(t) c^+ |gs> on ``linspace(start, stop, num)`` (reference:
    cmpy/exactdiag.py:248-273; signless ladder operators as there).  ``gs`` is an ``EigenState``."""
    n_up, n_dn = gs.n_up, gs.n_dn
    sector = model.basis.get_sector(n_up, n_dn)
    logger.debug("Computing greater GF (Sector: %d, %d; num: %d)", n_up, n_dn, num)
    sector_p1 = model.basis.upper_sector(n_up, n_dn, sigma)
    if sector_p1 is None:
        logger.warning("Upper sector not found!")
        times = np.linspace(start, stop, num)
        return times, np.zeros_like(times)
    return _autocorrelation(model, gs, sector_p1, CreationOperator, sector, start, stop, num, pos, sigma, +1)


def gf_lesser(model, gs, start, stop, num=1000, pos=0, sigma=UP):
    """G^<(t) = +i <gs| c^+ c(t) |gs> (reference: cmpy/exactdiag.py:276-301)."""
    n_up, n_dn = gs.n_up, gs.n_dn
    sector = model.basis.get_sector(n_up, n_dn)
    logger.debug("Computing lesser GF (Sector: %d, %d; num: %d)", n_up, n_dn, num)
    sector_m1 = model.basis.lower_sector(n_up, n_dn, sigma)
    if sector_m1 is None:
        logger.warning("Lower sector not found!")
        times = np.linspace(start, stop, num)
        return times, np.zeros_like(times)
    return _autocorrelation(model, gs, sector_m1, AnnihilationOperator, sector, start, stop, num, pos, sigma, -1)


def gf_tevo(model, start, stop, num=1000, pos=0, sigma=UP):
    """G^>(t) - G^<(t) from the ground state over all sectors (reference: cmpy/exactdiag.py:304-308)."""
    gs = compute_groundstate(model)
    times, gf_g = gf_greater(model, gs, start, stop, num, pos, sigma)
    times, gf_l = gf_lesser(model, gs, start, stop, num, pos, sigma)
    return times, gf_g - gf_l


def fourier_t2z(times, gf_t, omegas, delta=1e-2, eta=None):
    """Laplace transform G(z) = int dt exp(i z t) G(t), z = omegas + i eta (reference:
    cmpy/exactdiag.py:311-316, which calls gftool.fourier.tt2z; gftool >= 0.10 is an unpinned,
    un-vendored dependency that is absent here, so this restates its default piecewise-linear
    rule `tt2z_lin` -- PARITY UNPINNED for this one function).  Evaluated on the device."""
    torch = _lib.require_cuda()
    times = np.asarray(times, dtype=np.float64)
    if eta is None:
        eta = -np.log(delta) / times[-1]
    z = np.asarray(omegas) + 1j * eta
    dev = _lib.device()
    t = torch.from_numpy(times).to(dev)
    g = torch.from_numpy(np.asarray(gf_t, dtype=np.complex128)).to(dev)
    zt = torch.from_numpy(np.ascontiguousarray(z, dtype=np.complex128).reshape(-1)).to(dev)
    dt = t[1:] - t[:-1]
    a = 1j * zt[:, None]                                   # (nz, 1)
    ea = torch.exp(a * dt[None, :])                        # exp(i z dt_n)
    e0 = torch.exp(a * t[None, :-1])                       # exp(i z t_n)
    dg = (g[1:] - g[:-1])[None, :]
    # int_0^dt exp(a s) (g_n + dg s/dt) ds = g_n (ea-1)/a + dg/dt (dt ea / a - (ea-1)/a^2)
    term = g[None, :-1] * (ea - 1) / a + dg / dt[None, :] * (dt[None, :] * ea / a - (ea - 1) / a ** 2)
    gz = (e0 * term).sum(dim=1)
    return z, gz.cpu().numpy().reshape(np.shape(z))
