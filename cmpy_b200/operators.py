# -*- coding: utf-8 -*-
"""Linear operators and Hamiltonian projectors with cmpy's interface, backed by the
sm_100a kernels of ``libcmpy_b200.so`` (reference: cmpy/operators.py).

* ``project_onsite_energy / project_hubbard_inter / project_hopping`` keep the reference's
  names, argument order (including ``num_sites`` = fermion-sign width) and *emission order*;
  the per-string energies, hop targets and signs they iterate over are built on the GPU
  (kernels K2/K3, csrc/sector.cuh).
* ``HamiltonOperator(size, data, indices)`` stays source compatible (COO input, GPU COO
  mat-vec).  ``SectorHamiltonOperator`` / ``SpinHamiltonOperator`` are the matrix-free
  operators the models return: no triplets are ever materialised for H.v.
* ``_matvec`` accepts numpy vectors (host round trip) and, through ``matvec``/``apply``,
  torch CUDA tensors (zero copy).
"""
import abc
import ctypes
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np
import scipy.sparse.linalg as sla
from scipy.sparse import csr_matrix

from . import _lib
from .basis import UP, SPIN_CHARS

__all__ = [
    "LinearOperator", "HamiltonOperator", "SectorHamiltonOperator", "SpinHamiltonOperator",
    "CreationOperator", "AnnihilationOperator", "project_up", "project_dn",
    "project_elements_up", "project_elements_dn", "project_hubbard_inter",
    "project_onsite_energy", "project_hopping", "TimeEvolutionOperator",
    "species_hops", "weighted_elements",
]


# =========================================================================================
# index projections (pure index arithmetic, reference: cmpy/operators.py:33-90)
# =========================================================================================

def project_up(up_idx, num_dn_states, dn_indices):
    """Indices of the up-string ``up_idx`` combined with the given dn indices."""
    return np.atleast_1d(up_idx * num_dn_states + dn_indices)


def project_dn(dn_idx, num_dn_states, up_indices):
    """Indices of the dn-string ``dn_idx`` combined with the given up indices."""
    return np.atleast_1d(up_indices * num_dn_states + dn_idx)


def project_elements_up(up_idx, num_dn_states, dn_indices, value, target=None):
    """Yields ``(row, col, value)`` for an up-string element spread over the dn indices
    (reference: cmpy/operators.py:93-155)."""
    rows = project_up(up_idx, num_dn_states, dn_indices)
    cols = rows if target is None else project_up(target, num_dn_states, dn_indices)
    for row, col in zip(rows, cols):
        yield row, col, value


def project_elements_dn(dn_idx, num_dn_states, up_indices, value, target=None):
    """Yields ``(row, col, value)`` for a dn-string element spread over the up indices
    (reference: cmpy/operators.py:158-220)."""
    rows = project_dn(dn_idx, num_dn_states, up_indices)
    cols = rows if target is None else project_dn(target, num_dn_states, up_indices)
    for row, col in zip(rows, cols):
        yield row, col, value


# =========================================================================================
# GPU-built per-string tables (K2 / K3)
# =========================================================================================

def _states_tensor(states):
    torch = _lib.require_cuda()
    if isinstance(states, torch.Tensor):
        return states.to(device=_lib.device(), dtype=torch.int64).contiguous()
    arr = np.ascontiguousarray(np.asarray(states, dtype=np.int64))
    return torch.from_numpy(arr).to(_lib.device())


def is_full_sector(states) -> bool:
    """True when ``states`` is the complete ascending list of fixed-popcount integers (then
    the combinadic rank equals the reference's ``bisect_left``)."""
    arr = np.asarray(states, dtype=np.int64)
    if arr.ndim != 1 or arr.size == 0:
        return False
    top = int(arr[-1])
    n = int(arr[0]).bit_count()
    if top.bit_count() != n:
        return False
    width = top.bit_length()
    if width > 62 or arr.size != _lib.binomial(width, n):
        return False
    # complete iff first/last are the extreme strings and the length matches
    return int(arr[0]) == (1 << n) - 1 and top == ((1 << n) - 1) << (width - n)


def species_hops(states, width, site1, site2, fixed_popcount=None):
    """Hop table of one species for the bond ``site1 < site2`` (kernel K2).

    Returns ``(target, sign)`` numpy arrays: ``target[i]`` is the index of the string reached
    from ``states[i]`` (``-1`` if the two sites are equally occupied), ``sign[i]`` the fermion
    sign limited to bits ``< width`` (reference: cmpy/operators.py:253-273, 425-460)."""
    assert site1 < site2  # reference: cmpy/operators.py:438
    torch = _lib.require_cuda()
    st = _states_tensor(states)
    if fixed_popcount is None:
        fixed_popcount = is_full_sector(st.cpu().numpy())
    tgt = torch.empty(st.numel(), dtype=torch.int32, device=st.device)
    sgn = torch.empty(st.numel(), dtype=torch.int8, device=st.device)
    _lib.check(_lib.lib().cmpy_species_hops(_lib.ptr(st), st.numel(), int(bool(fixed_popcount)),
                                            int(width), int(site1), int(site2), _lib.ptr(tgt),
                                            _lib.ptr(sgn), _lib.stream_ptr()), "cmpy_species_hops")
    return tgt.cpu().numpy(), sgn.cpu().numpy()


def weighted_elements(states, values) -> np.ndarray:
    """``weighted_element(state, values)`` for every state (kernel K3; reference:
    cmpy/operators.py:226-250 -- ascending-site summation order, bit exact)."""
    torch = _lib.require_cuda()
    st = _states_tensor(states)
    vals, vals_p = _lib.as_c_array(np.atleast_1d(values), c_double, np.float64)
    out = torch.empty(st.numel(), dtype=torch.float64, device=st.device)
    _lib.check(_lib.lib().cmpy_weighted_elements(_lib.ptr(st), st.numel(), vals_p, len(vals),
                                                 _lib.ptr(out), _lib.stream_ptr()),
               "cmpy_weighted_elements")
    return out.cpu().numpy()


def inter_elements(up_states, dn_states, u) -> np.ndarray:
    """Interaction energy of every (up, dn) pair, up-major (reference:
    cmpy/operators.py:305-356)."""
    torch = _lib.require_cuda()
    up, dn = _states_tensor(up_states), _states_tensor(dn_states)
    vals, vals_p = _lib.as_c_array(np.atleast_1d(u), c_double, np.float64)
    out = torch.empty(up.numel() * dn.numel(), dtype=torch.float64, device=up.device)
    _lib.check(_lib.lib().cmpy_inter_elements(_lib.ptr(up), up.numel(), _lib.ptr(dn), dn.numel(),
                                              vals_p, len(vals), _lib.ptr(out), _lib.stream_ptr()),
               "cmpy_inter_elements")
    return out.cpu().numpy()


# =========================================================================================
# Hamiltonian projectors (generators of COO triplets, reference emission order)
# =========================================================================================

def project_hubbard_inter(up_states, dn_states, u):
    """Yields ``(idx, idx, sum_i u[i] [up & dn]_i)`` up-major, zero energies skipped
    (reference: cmpy/operators.py:305-356)."""
    energies = inter_elements(up_states, dn_states, u)
    for origin in np.nonzero(energies)[0]:
        yield int(origin), int(origin), float(energies[origin])


def project_onsite_energy(up_states, dn_states, eps):
    """Yields the on-site energy triplets: all spin-up entries (each up string spread over
    every dn index), then all spin-down entries; zero energies skipped
    (reference: cmpy/operators.py:359-422)."""
    num_up, num_dn = len(up_states), len(dn_states)
    e_up = weighted_elements(up_states, eps)
    e_dn = weighted_elements(dn_states, eps)
    for up_idx in np.nonzero(e_up)[0]:
        energy = float(e_up[up_idx])
        base = int(up_idx) * num_dn
        for origin in range(base, base + num_dn):
            yield origin, origin, energy
    for dn_idx in np.nonzero(e_dn)[0]:
        energy = float(e_dn[dn_idx])
        for origin in range(int(dn_idx), num_up * num_dn, num_dn):
            yield origin, origin, energy


def _compute_hopping_term(states, width, site1, site2, hop):
    """Yields ``(origin, target, sign*hop)`` for one species
    (reference: cmpy/operators.py:436-460)."""
    tgt, sgn = species_hops(states, width, site1, site2)
    for i in np.nonzero(tgt >= 0)[0]:
        yield int(i), int(tgt[i]), int(sgn[i]) * hop


def project_hopping(up_states, dn_states, num_sites, site1, site2, hop):
    """Yields the hopping triplets of the bond ``site1 < site2``: the spin-up block (every
    up hop spread over all dn indices) followed by the spin-down block.  ``num_sites`` is the
    fermion-sign width (reference: cmpy/operators.py:463-527)."""
    num_up, num_dn = len(up_states), len(dn_states)
    for o, t, a in _compute_hopping_term(up_states, num_sites, site1, site2, hop):
        for d in range(num_dn):
            yield o * num_dn + d, t * num_dn + d, a
    for o, t, a in _compute_hopping_term(dn_states, num_sites, site1, site2, hop):
        for u in range(num_up):
            yield u * num_dn + o, u * num_dn + t, a


# =========================================================================================
# Linear operators
# =========================================================================================

class LinearOperator(sla.LinearOperator, abc.ABC):
    """scipy ``LinearOperator`` with ``toarray()``, ``trace()`` and scalar products that
    keep those methods (reference: cmpy/operators.py:535-608)."""

    def __init__(self, shape, dtype=None):
        sla.LinearOperator.__init__(self, shape=shape, dtype=dtype)
        abc.ABC.__init__(self)

    def __repr__(self):
        return f"{self.__class__.__name__}(shape: {self.shape}, dtype: {self.dtype})"

    def toarray(self) -> np.ndarray:
        return self.matmat(np.eye(self.shape[1], dtype=self.dtype))

    def _trace(self) -> float:
        return float(np.trace(self.toarray()))

    def trace(self) -> float:
        return self._trace()

    def _decorate(self, scaled, x):
        try:
            scaled.trace = lambda: x * self.trace()
            scaled.toarray = lambda: x * self.toarray()
        except AttributeError:
            pass
        return scaled

    def __mul__(self, x):
        return self._decorate(super().__mul__(x), x)

    def __rmul__(self, x):
        return self._decorate(super().__rmul__(x), x)


def pipelined_host_batch(owner, n, apply_fn, xs, outs=None):
    """Shared by the single-GPU and the sharded operator: ``apply_fn(dx, dy)`` computes dy = H dx on the
    current stream for device vectors of ``n`` elements; host vectors stream through two device buffers
    on three streams (H2D of vector i+1 || H.v of vector i || D2H of result i-1)."""
    torch = _lib.require_cuda()
    dev = _lib.device()
    st = getattr(owner, "_batch_state", None)
    if st is None:
        st = dict(h2d=torch.cuda.Stream(), comp=torch.cuda.Stream(), d2h=torch.cuda.Stream(),
                  dx=[torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(2)],
                  dy=[torch.empty(max(n, 1), dtype=torch.float64, device=dev) for _ in range(2)], outs=None)
        owner._batch_state = st
    if outs is None:
        if st["outs"] is None:
            st["outs"] = [torch.empty(max(n, 1), dtype=torch.float64).pin_memory() for _ in range(2)]
        outs = st["outs"]
    cur = torch.cuda.current_stream()
    for s_ in (st["h2d"], st["comp"], st["d2h"]):
        s_.wait_stream(cur)
    comp_done = [None, None]
    d2h_done = [None, None]
    results = []
    for i, x in enumerate(xs):
        if x.is_cuda or x.dtype != torch.float64 or x.numel() != n:
            raise TypeError("matvec_batch expects CPU float64 tensors of the operator's (local) size")
        b = i & 1
        if comp_done[b] is not None:
            st["h2d"].wait_event(comp_done[b])        # dx[b] was consumed by H.v i-2
        with torch.cuda.stream(st["h2d"]):
            st["dx"][b][:n].copy_(x.view(-1), non_blocking=True)
            ev_in = torch.cuda.Event(); ev_in.record()
        st["comp"].wait_event(ev_in)
        if d2h_done[b] is not None:
            st["comp"].wait_event(d2h_done[b])         # dy[b] was copied out (result i-2)
        with torch.cuda.stream(st["comp"]):
            apply_fn(st["dx"][b][:n], st["dy"][b][:n])
            comp_done[b] = torch.cuda.Event(); comp_done[b].record()
        st["d2h"].wait_event(comp_done[b])
        out = outs[i % len(outs)]
        with torch.cuda.stream(st["d2h"]):
            out.view(-1)[:n].copy_(st["dy"][b][:n], non_blocking=True)
            d2h_done[b] = torch.cuda.Event(); d2h_done[b].record()
        results.append(out)
    for s_ in (st["h2d"], st["comp"], st["d2h"]):
        s_.synchronize()
    return results


class _DeviceOperatorMixin:
    """Shared GPU plumbing: the C handle, host<->device staging, Lanczos entry."""

    _handle = None

    def _set_handle(self, handle):
        self._handle = handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.lib().cmpy_op_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        if not self._handle:
            raise RuntimeError("operator has no device handle")
        return self._handle

    def set_variant(self, variant: int):
        """0 = auto, 1 = global-gather kernel, 2 = shared-memory row kernel."""
        _lib.check(_lib.lib().cmpy_hv_set_variant(self.handle, int(variant)), "cmpy_hv_set_variant")

    def apply(self, x, out=None):
        """y = H x on the device. ``x``: CUDA float64 tensor of length ``shape[1]``."""
        torch = _lib.torch_mod()
        if x.dtype != torch.float64 or not x.is_cuda:
            raise TypeError("apply() expects a CUDA float64 tensor")
        x = x.contiguous().view(-1)
        if x.numel() != self.shape[1]:
            raise ValueError(f"dimension mismatch: {x.numel()} != {self.shape[1]}")
        if out is None:
            out = torch.empty(self.shape[0], dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().cmpy_hv_apply(self.handle, _lib.ptr(x), _lib.ptr(out),
                                            _lib.stream_ptr()), "cmpy_hv_apply")
        return out

    def _host_staging(self, n):
        """Two cached pinned host buffers of the operator's size (input and result staging)."""
        torch = _lib.require_cuda()
        st = getattr(self, "_pinned_io", None)
        if st is None or st[0].numel() != n:
            st = (torch.empty(n, dtype=torch.float64).pin_memory(),
                  torch.empty(n, dtype=torch.float64).pin_memory())
            self._pinned_io = st
        return st

    def _apply_any(self, x, out=None):
        """numpy in -> numpy out (host round trip); CUDA tensor in -> CUDA tensor out.

        Host data goes through two cached pinned staging buffers (multi-threaded host copy into
        the pinned input, asynchronous H2D, H.v, asynchronous D2H into the pinned result).  The
        returned host array / tensor is FRESH unless the caller passes ``out`` (a CPU float64
        tensor, pinned for full speed) -- the reference's ``_matvec`` never aliases its results
        (cmpy/operators.py:626-630)."""
        torch = _lib.require_cuda()
        if isinstance(x, torch.Tensor) and not x.is_cuda and not x.is_complex():
            n = x.numel()
            src = x.contiguous().view(-1)
            if src.dtype != torch.float64 or not src.is_pinned():
                pin_in, _ = self._host_staging(n)
                pin_in.copy_(src)
                src = pin_in
            dev = src.to(device=_lib.device(), non_blocking=True)
            y = self.apply(dev)
            if out is not None:
                out.view(-1).copy_(y, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return out.view(x.shape)
            _, pin_out = self._host_staging(n)
            pin_out.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return pin_out.clone().view(x.shape)
        if isinstance(x, torch.Tensor):
            if x.is_complex():
                xr = torch.view_as_real(x.contiguous().view(-1))
                y = torch.complex(self.apply(xr[:, 0].contiguous()), self.apply(xr[:, 1].contiguous()))
                return y
            return self.apply(x.to(device=_lib.device(), dtype=torch.float64))
        arr = np.asarray(x)
        shape = arr.shape
        flat = arr.reshape(-1)
        if np.iscomplexobj(flat):
            re = self._apply_any(np.ascontiguousarray(flat.real))
            im = self._apply_any(np.ascontiguousarray(flat.imag))
            return (re + 1j * im).astype(np.result_type(flat.dtype, np.complex128)).reshape(shape)
        # what scipy (eigsh, expm_multiply) hands over: a pageable numpy vector
        host = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64))
        # fresh result from numpy's allocator (large arrays are madvise(HUGEPAGE)d: far fewer first-touch faults
        # than torch.empty -- measured 128 vs 229 ms for the copy into a fresh 1.3 GB array)
        res_np = np.empty(host.numel(), dtype=np.float64)
        self._pageable_roundtrip(host, torch.from_numpy(res_np))
        return res_np.reshape(shape)

    def _pageable_roundtrip(self, host, res, chunks=8):
        """``res = H host`` for pageable host vectors, in ``chunks`` pieces through the pinned staging buffers:
        the host copy of piece k+1 into the pinned input overlaps the H2D copy of piece k, and the copy of result
        piece k out of the pinned buffer overlaps the D2H copy of piece k+1 (the host copies, not PCIe, are the
        longer leg).  One H.v on the whole vector in between."""
        torch = _lib.require_cuda()
        n = host.numel()
        pin_in, pin_out = self._host_staging(n)
        dev = getattr(self, "_dev_io", None)
        if dev is None or dev[0].numel() != n:
            dev = (torch.empty(n, dtype=torch.float64, device=_lib.device()),
                   torch.empty(n, dtype=torch.float64, device=_lib.device()))
            self._dev_io = dev
        dx, dy = dev
        cur = torch.cuda.current_stream()
        side = getattr(self, "_io_stream", None)
        if side is None:
            side = self._io_stream = torch.cuda.Stream()
        chunks = max(1, min(int(chunks), n // (1 << 20) or 1))
        bounds = [(n * k) // chunks for k in range(chunks + 1)]
        side.wait_stream(cur)               # dx / dy are free (earlier calls on the current stream are done with them)
        with torch.cuda.stream(side):
            for a, b in zip(bounds[:-1], bounds[1:]):
                pin_in[a:b].copy_(host[a:b])                    # host copy (blocking), then the asynchronous H2D
                dx[a:b].copy_(pin_in[a:b], non_blocking=True)
        cur.wait_stream(side)
        self.apply(dx, out=dy)
        side.wait_stream(cur)
        events = []
        with torch.cuda.stream(side):
            for a, b in zip(bounds[:-1], bounds[1:]):
                pin_out[a:b].copy_(dy[a:b], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(side)
                events.append(ev)
        for ev, a, b in zip(events, bounds[:-1], bounds[1:]):
            ev.synchronize()
            res[a:b].copy_(pin_out[a:b])
        cur.wait_stream(side)
        return res

    def matvec(self, x, out=None):
        torch = _lib.torch_mod()
        if isinstance(x, torch.Tensor):
            return self._apply_any(x, out=out)
        return super().matvec(x)

    def matvec_batch(self, xs, outs=None):
        """H applied to a sequence of HOST vectors (the column-by-column ``matmat`` of the reference,
        cmpy/exactdiag.py:132-134), pipelined: the H2D copy of vector i+1, the H.v of vector i and
        the D2H copy of result i-1 run on three streams with double-buffered device vectors, so a
        long batch is bound by one PCIe direction instead of the sum of both.

        ``xs``: CPU float64 tensors (pinned memory for full speed).  ``outs``: optional CPU tensors
        receiving the results (cycled if shorter than ``xs``); default two pinned buffers, i.e. only
        the last two results stay valid.  Returns the list of output tensors, one per input."""
        return pipelined_host_batch(self, self.shape[0], lambda dx, dy: self.apply(dx, out=dy), xs, outs)

    def diagonal(self) -> np.ndarray:
        torch = _lib.require_cuda()
        d = torch.empty(self.shape[0], dtype=torch.float64, device=_lib.device())
        _lib.check(_lib.lib().cmpy_op_diagonal(self.handle, _lib.ptr(d), _lib.stream_ptr()),
                   "cmpy_op_diagonal")
        return d.cpu().numpy()

    def _device_trace(self) -> float:
        out = c_double(0.0)
        _lib.check(_lib.lib().cmpy_op_trace(self.handle, ctypes.byref(out)), "cmpy_op_trace")
        return float(out.value)

    def _dense_from_device(self) -> np.ndarray:
        """Dense matrix by applying H to unit vectors on the device."""
        torch = _lib.require_cuda()
        n = self.shape[0]
        eye = torch.eye(n, dtype=torch.float64, device=_lib.device())
        out = torch.empty_like(eye)
        for k in range(n):
            self.apply(eye[k], out=out[k])
        return out.cpu().numpy().T.copy()


class HamiltonOperator(_DeviceOperatorMixin, LinearOperator):
    """Hamiltonian as LinearOperator built from COO triplets (source compatible with
    cmpy/operators.py:614-646): ``y[col] += val * x[row]`` on the GPU."""

    def __init__(self, size, data, indices, dtype=None):
        data = np.asanyarray(data)
        indices = np.asanyarray(indices)
        if dtype is None:
            dtype = data.dtype
        super().__init__((size, size), dtype=dtype)
        self.data = data
        self.indices = indices.T  # (nnz, 2) rows of (row, col), as the reference keeps them
        self._build_coo()

    def _build_coo(self):
        _lib.require_cuda()
        idx = np.asarray(self.indices).reshape(-1, 2) if self.indices.size else np.zeros((0, 2), np.int64)
        rows, rows_p = _lib.as_c_array(idx[:, 0], c_int64, np.int64)
        cols, cols_p = _lib.as_c_array(idx[:, 1], c_int64, np.int64)
        if np.iscomplexobj(self.data):
            raise TypeError("complex COO data is not supported by the GPU HamiltonOperator")
        vals, vals_p = _lib.as_c_array(self.data.reshape(-1), c_double, np.float64)
        handle = c_void_p()
        _lib.check(_lib.lib().cmpy_coo_create(int(self.shape[0]), len(vals), rows_p, cols_p, vals_p,
                                              ctypes.byref(handle)), "cmpy_coo_create")
        self._set_handle(handle)

    def _matvec(self, x) -> np.ndarray:
        return self._apply_any(x)

    def toarray(self):
        idx = self.indices.reshape(-1, 2)
        csr = csr_matrix((self.data, (idx[:, 0], idx[:, 1])), shape=self.shape, dtype=self.dtype)
        return csr.toarray()

    def _adjoint(self) -> "HamiltonOperator":
        """Hamiltonian is hermitian."""
        return self

    def _trace(self) -> float:
        return self._device_trace()


class SectorHamiltonOperator(HamiltonOperator):
    """Matrix-free Hubbard / Anderson Hamiltonian on one (n_up, n_dn) sector (kernel K4).

    Built by ``AbstractManyBodyModel.hamilton_operator``; never materialises triplets for
    the mat-vec.  ``data`` / ``indices`` are produced lazily (reference emission order) only
    if somebody asks for them."""

    def __init__(self, num_sites, up_states, dn_states, bonds, hops, eps, u, sign_width,
                 dtype=None, fixed_popcount=None):
        up = np.ascontiguousarray(np.asarray(up_states, dtype=np.int64))
        dn = np.ascontiguousarray(np.asarray(dn_states, dtype=np.int64))
        size = len(up) * len(dn)
        LinearOperator.__init__(self, (size, size), dtype=dtype or np.float64)
        self.num_sites = int(num_sites)
        self.up_states, self.dn_states = up, dn
        self.bonds = [(int(i), int(j)) for i, j in bonds]
        self.hops = np.asarray(hops, dtype=np.float64).reshape(-1)
        self.eps = np.asarray(eps, dtype=np.float64).reshape(-1)
        self.u = np.asarray(u, dtype=np.float64).reshape(-1)
        self.sign_width = int(sign_width)
        self._coo = None
        for i, j in self.bonds:
            assert i < j  # reference: cmpy/operators.py:438
        if fixed_popcount is None:
            fixed_popcount = is_full_sector(up) and is_full_sector(dn)
        _lib.require_cuda()
        up_p = up.ctypes.data_as(POINTER(c_int64))
        dn_p = dn.ctypes.data_as(POINTER(c_int64))
        b_arr, b_p = _lib.as_c_array(np.asarray(self.bonds, dtype=np.int32).reshape(-1), c_int32, np.int32)
        h_arr, h_p = _lib.as_c_array(self.hops, c_double, np.float64)
        e_arr, e_p = _lib.as_c_array(self.eps, c_double, np.float64)
        u_arr, u_p = _lib.as_c_array(self.u, c_double, np.float64)
        if len(e_arr) != self.num_sites or len(u_arr) != self.num_sites or len(h_arr) != len(self.bonds):
            raise ValueError("eps/u need num_sites entries and hops one entry per bond")
        handle = c_void_p()
        _lib.check(_lib.lib().cmpy_hubbard_create(
            self.num_sites, up_p, len(up), dn_p, len(dn), int(bool(fixed_popcount)), len(self.bonds),
            b_p, h_p, e_p, u_p, self.sign_width, ctypes.byref(handle)), "cmpy_hubbard_create")
        self._set_handle(handle)

    # -- lazy COO view (compatibility with code that reads .data / .indices) -------------
    def _triplets(self):
        if self._coo is None:
            rows, cols, vals = [], [], []
            gens = [project_onsite_energy(self.up_states, self.dn_states, self.eps),
                    project_hubbard_inter(self.up_states, self.dn_states, self.u)]
            for (i, j), t in zip(self.bonds, self.hops):
                gens.append(project_hopping(self.up_states, self.dn_states, self.sign_width, i, j, t))
            for gen in gens:
                for r, c, v in gen:
                    rows.append(r); cols.append(c); vals.append(v)
            self._coo = (np.asarray(vals, dtype=np.float64),
                         np.stack([np.asarray(rows, np.int64), np.asarray(cols, np.int64)], axis=1)
                         if rows else np.zeros((0, 2), np.int64))
        return self._coo

    @property
    def data(self):
        return self._triplets()[0]

    @data.setter
    def data(self, value):  # assigned by nobody; kept so attribute assignment does not fail
        pass

    @property
    def indices(self):
        return self._triplets()[1]

    @indices.setter
    def indices(self, value):
        pass

    def apply_rows(self, x_slab, row0, nrows, out=None, accumulate=False):
        """(D + dn hops) on the slab of up-rows [row0, row0+nrows) -- local phase of the
        up-string-sharded H.v."""
        torch = _lib.torch_mod()
        num_dn = len(self.dn_states)
        if out is None:
            out = torch.empty(nrows * num_dn, dtype=torch.float64, device=x_slab.device)
        _lib.check(_lib.lib().cmpy_hubbard_apply_rows(self.handle, _lib.ptr(x_slab), _lib.ptr(out),
                                                      int(row0), int(nrows), int(accumulate),
                                                      _lib.stream_ptr()), "cmpy_hubbard_apply_rows")
        return out

    def toarray(self):
        return self._dense_from_device().astype(self.dtype, copy=False)


class SpinHamiltonOperator(HamiltonOperator):
    """Matrix-free Heisenberg / XXZ Hamiltonian on a magnetisation sector (kernel K5)."""

    def __init__(self, num_sites, n_up, pairs, j, jz, dtype=None):
        self.num_sites = int(num_sites)
        self.n_up = -1 if n_up is None else int(n_up)
        self.pairs = [(int(a), int(b)) for a, b in pairs]
        self.j, self.jz = float(j), float(jz)
        _lib.require_cuda()
        p_arr, p_p = _lib.as_c_array(np.asarray(self.pairs, dtype=np.int32).reshape(-1), c_int32, np.int32)
        handle = c_void_p()
        _lib.check(_lib.lib().cmpy_heisenberg_create(self.num_sites, self.n_up, len(self.pairs), p_p,
                                                     self.j, self.jz, ctypes.byref(handle)),
                   "cmpy_heisenberg_create")
        size = c_int64(0)
        _lib.check(_lib.lib().cmpy_op_size(handle, ctypes.byref(size)), "cmpy_op_size")
        LinearOperator.__init__(self, (size.value, size.value), dtype=dtype or np.float64)
        self._set_handle(handle)

    @property
    def data(self):
        raise AttributeError("matrix-free operator: use the model's hamiltonian_data()")

    @data.setter
    def data(self, value):
        pass

    @property
    def indices(self):
        raise AttributeError("matrix-free operator: use the model's hamiltonian_data()")

    @indices.setter
    def indices(self, value):
        pass

    def toarray(self):
        return self._dense_from_device().astype(self.dtype, copy=False)


# =========================================================================================
# Creation / annihilation operators (kernel K6)
# =========================================================================================

class _LadderOperator(LinearOperator):
    dagger = True

    def __init__(self, sector, sector_t, pos=0, sigma=UP, signed=False):
        dim_origin = sector.size
        if sigma == UP:
            dim_target = sector_t.num_up * sector.num_dn
        else:
            dim_target = sector_t.num_dn * sector.num_up
        # declared dtype complex64 as in the reference (cmpy/operators.py:716,760)
        super().__init__(shape=(dim_target, dim_origin), dtype=np.complex64)
        self.pos = pos
        self.sigma = sigma
        self.sector = sector
        self.signed = bool(signed)
        self._sector_t = sector_t
        self._dev = None

    def __repr__(self):
        name = f"{self.__class__.__name__}_{self.pos}{SPIN_CHARS[self.sigma]}"
        return f"{name}(shape: {self.shape}, dtype: {self.dtype})"

    def _device_lists(self):
        if self._dev is None:
            s, t = self.sector, self._sector_t
            up, dn = _states_tensor(s.up_states), _states_tensor(s.dn_states)
            if self.sigma == UP:
                up_t, dn_t = _states_tensor(t.up_states), dn
            else:
                up_t, dn_t = up, _states_tensor(t.dn_states)
            self._dev = (up, dn, up_t, dn_t)
        return self._dev

    def apply(self, x, out=None):
        """Device apply: ``x`` CUDA float64 (or complex128) tensor of the origin sector."""
        torch = _lib.require_cuda()
        up, dn, up_t, dn_t = self._device_lists()
        cplx = x.is_complex()
        xv = torch.view_as_real(x.contiguous()).contiguous() if cplx else x.contiguous()
        if out is None:
            out = torch.empty(self.shape[0], dtype=x.dtype, device=x.device)
        ov = torch.view_as_real(out) if cplx else out
        _lib.check(_lib.lib().cmpy_ladder_apply(
            _lib.ptr(up), up.numel(), _lib.ptr(dn), dn.numel(), _lib.ptr(up_t), up_t.numel(),
            _lib.ptr(dn_t), dn_t.numel(), int(self.pos), int(self.sigma), int(self.dagger),
            int(self.signed), 2 if cplx else 1, _lib.ptr(xv), _lib.ptr(ov), _lib.stream_ptr()),
            "cmpy_ladder_apply")
        return out

    def _matvec(self, x):
        torch = _lib.require_cuda()
        arr = np.asarray(x)
        tail = arr.shape[1:]
        flat = arr.reshape(arr.shape[0], -1)
        cols = []
        for k in range(flat.shape[1]):
            col = np.ascontiguousarray(flat[:, k])
            if np.iscomplexobj(col):
                dev = torch.from_numpy(col.astype(np.complex128)).to(_lib.device())
            else:
                dev = torch.from_numpy(col.astype(np.float64)).to(_lib.device())
            cols.append(self.apply(dev).cpu().numpy().astype(arr.dtype, copy=False))
        out = np.stack(cols, axis=1) if cols else np.zeros((self.shape[0], 0), dtype=arr.dtype)
        return out.reshape((self.shape[0], *tail))

    def matvec(self, x):
        torch = _lib.torch_mod()
        if isinstance(x, torch.Tensor):
            return self.apply(x)
        return super().matvec(x)


class CreationOperator(_LadderOperator):
    """Fermionic creation operator ``c^dagger_{pos,sigma}`` from ``sector`` to ``sector_p1``.

    ``signed=False`` (default) reproduces the reference: amplitudes are copied without a
    fermionic sign (cmpy/operators.py:652-676).  For ``sigma=DN`` the target row stride is
    the *target* sector's ``num_dn`` -- the reference uses the origin's and fails with an
    IndexError (cmpy/operators.py:668,675); this implementation is the corrected one."""

    dagger = True

    def __init__(self, sector, sector_p1, pos=0, sigma=UP, signed=False):
        super().__init__(sector, sector_p1, pos, sigma, signed)
        self.sector_p1 = sector_p1

    def _adjoint(self):
        return AnnihilationOperator(self.sector_p1, self.sector, self.pos, self.sigma, self.signed)


class AnnihilationOperator(_LadderOperator):
    """Fermionic annihilation operator ``c_{pos,sigma}`` from ``sector`` to ``sector_m1``
    (reference: cmpy/operators.py:679-703, 750-791)."""

    dagger = False

    def __init__(self, sector, sector_m1, pos=0, sigma=UP, signed=False):
        super().__init__(sector, sector_m1, pos, sigma, signed)
        self.sector_m1 = sector_m1

    def _adjoint(self):
        return CreationOperator(self.sector_m1, self.sector, self.pos, self.sigma, self.signed)


class TimeEvolutionOperator(LinearOperator):
    """Out of scope of the B200 hot path (dense eigendecomposition, reference:
    cmpy/operators.py:797-833); kept as a name so imports do not break."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "TimeEvolutionOperator is a dense O(N^2) utility outside the accelerated hot path "
            "(SURVEY.md section 2, component 5)")
