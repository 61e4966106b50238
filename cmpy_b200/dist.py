# -*- coding: utf-8 -*-
"""Up-string-sharded Hubbard H.v across the GPUs of one box.

The amplitude matrix X (num_up x num_dn, row-major: idx = up_idx*num_dn + dn_idx, reference
cmpy/operators.py:33-90) is split into contiguous slabs of up-rows, one slab per rank
(one process per GPU, ``torch.distributed``).  Diagonal and dn-hops act inside a row, so
they are local; up-hops couple rows, so they are applied in the transposed (dn-major)
layout, where they are again row-local:

  1. y_local  = (D + T_dn) x_local                         local kernel (slab of up-rows)
  2. pack     : X[R_p, C_q] -> (C_q x R_p) blocks          transpose kernel, one block per peer
  3. exchange : all-to-all (NCCL, grouped send/recv)       -> rank q holds X^T[C_q, :]
  4. place    : blocks -> XT_local (|C_q| x num_up)        pitched copy kernel
  5. YT_local = T_up XT_local                              same local kernel, roles swapped
  6. pack / exchange / accumulate back into y_local        transpose + all-to-all + copy2d(+=)

On CUDA with more than one rank the default exchange is ``"peer"``: the XT / YT slabs live in
symmetric memory (``torch.distributed._symmetric_memory``), steps 2-4 are ONE kernel that reads
the local slab and stores the transposed tiles straight into the owners' XT slabs over NVLink
(``cmpy_transpose_push``), step 6 is ONE kernel that loads the peers' YT tiles over NVLink and
accumulates them into y (``cmpy_transpose_pull_acc``); cross-rank barriers order the phases and
the push runs on a side stream under the local dn pass.  No staging slabs: x, y, XT, YT only.
``exchange="a2a"`` keeps the NCCL all-to-all choreography (also what the gloo CPU tests drive).

Two transposes per H.v move 2 * 8 * dim * (P-1)/P bytes over NVLink.  The reference has no
distributed path at all; results are checked against the single-GPU operator / the oracle.

The exchange choreography is independent of the device kernels: ``LocalBackend`` objects
provide the four local primitives, so the same code runs under ``gloo`` on CPU in
tests/ with a checker backend (no product path uses that).
"""
import numpy as np

from . import _lib

__all__ = ["ShardPlan", "ShardedHubbardOperator", "CudaBackend", "lanczos_sharded",
           "gf_continued_fraction_sharded"]


class ShardPlan:
    """Balanced contiguous partition of the up-rows (and, for the transposed phase, of the
    dn-columns) over ``world`` ranks."""

    def __init__(self, num_up, num_dn, world, rank):
        self.num_up, self.num_dn, self.world, self.rank = int(num_up), int(num_dn), int(world), int(rank)
        self.row_bounds = [(self.num_up * k) // self.world for k in range(self.world + 1)]
        self.col_bounds = [(self.num_dn * k) // self.world for k in range(self.world + 1)]

    def rows(self, p=None):
        p = self.rank if p is None else p
        return self.row_bounds[p], self.row_bounds[p + 1]

    def cols(self, q=None):
        q = self.rank if q is None else q
        return self.col_bounds[q], self.col_bounds[q + 1]

    @property
    def nrows(self):
        r0, r1 = self.rows()
        return r1 - r0

    @property
    def ncols(self):
        c0, c1 = self.cols()
        return c1 - c0

    @property
    def local_size(self):
        return self.nrows * self.num_dn

    @property
    def local_size_t(self):
        return self.ncols * self.num_up

    def fwd_send_counts(self):
        """elements sent to each peer q in the forward transpose: |R_p| * |C_q|"""
        return [self.nrows * (self.cols(q)[1] - self.cols(q)[0]) for q in range(self.world)]

    def fwd_recv_counts(self):
        """elements received from each peer p: |C_q| * |R_p| (q = this rank)"""
        return [self.ncols * (self.rows(p)[1] - self.rows(p)[0]) for p in range(self.world)]

    def bytes_out_per_hv(self):
        """bytes leaving this GPU per H.v (two transposes, own block excluded)"""
        own = self.nrows * self.ncols
        return 8 * ((self.local_size - own) + (self.local_size_t - own))


class CudaBackend:
    """Local primitives on the device through the C ABI."""

    def __init__(self, op_main, op_t):
        self.op_main, self.op_t = op_main, op_t
        self.torch = _lib.require_cuda()

    def empty(self, n):
        return self.torch.empty(max(int(n), 1), dtype=self.torch.float64, device=_lib.device())

    def apply_rows(self, x, row0, nrows, out, accumulate=False):
        return self.op_main.apply_rows(x, row0, nrows, out=out, accumulate=accumulate)

    def apply_rows_t(self, xt, col0, ncols, out):
        return self.op_t.apply_rows(xt, col0, ncols, out=out)

    def transpose(self, src, src_off, nrows, ncols, ld_in, dst, dst_off, ld_out):
        esz = 8
        _lib.check(_lib.lib().cmpy_transpose(
            _lib.c_void_p(src.data_ptr() + esz * src_off), nrows, ncols, ld_in,
            _lib.c_void_p(dst.data_ptr() + esz * dst_off), ld_out, 0, _lib.stream_ptr()), "cmpy_transpose")

    def copy2d(self, src, src_off, nrows, ncols, ld_in, dst, dst_off, ld_out, accumulate):
        esz = 8
        _lib.check(_lib.lib().cmpy_copy2d(
            _lib.c_void_p(src.data_ptr() + esz * src_off), nrows, ncols, ld_in,
            _lib.c_void_p(dst.data_ptr() + esz * dst_off), ld_out, int(accumulate), _lib.stream_ptr()),
            "cmpy_copy2d")


class ShardedHubbardOperator:
    """H.v on the slab of up-rows owned by this rank; ``apply_local(x_local)`` returns
    ``(H x)_local``.  ``model`` is a ``HubbardModel`` / ``SingleImpurityAndersonModel``."""

    def __init__(self, model, n_up, n_dn, group=None, backend=None, up_states=None, dn_states=None,
                 exchange=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.model = model
        spec = model._operator_spec()
        if backend is None:
            from .operators import SectorHamiltonOperator

            sector = model.basis.get_sector(n_up, n_dn)
            up_states, dn_states = np.asarray(sector.up_states), np.asarray(sector.dn_states)
            nsites = model.num_sites
            op_main = SectorHamiltonOperator(nsites, up_states, dn_states, spec["bonds"], spec["hops"],
                                             spec["eps"], spec["u"], spec["sign_width"])
            zeros = np.zeros(nsites)
            # transposed phase: rows = dn strings, "dn hops" of that operator = up hops of H
            op_t = SectorHamiltonOperator(nsites, dn_states, up_states, spec["bonds"], spec["hops"],
                                          zeros, zeros, spec["sign_width"])
            backend = CudaBackend(op_main, op_t)
        self.backend = backend
        self._cdist = None
        self.plan = ShardPlan(len(up_states), len(dn_states), self.world, self.rank)
        size = self.plan.num_up * self.plan.num_dn
        self.shape = (size, size)
        p = self.plan
        if exchange is None:
            exchange = "peer" if (isinstance(backend, CudaBackend) and self.world > 1) else "a2a"
        if exchange not in ("peer", "a2a"):
            raise ValueError("exchange must be 'peer' or 'a2a'")
        self.exchange = exchange
        if exchange == "peer":
            try:
                self._init_peer()
                ok = 1
            except Exception as exc:  # no peer mapping on this box (no NVLink / no fabric handles)
                import logging

                logging.getLogger("cmpy").warning("peer-memory exchange unavailable (%s); using all-to-all", exc)
                ok = 0
            if self.world > 1:  # every rank must take the same path
                torch = _lib.require_cuda()
                flag = torch.tensor([ok], dtype=torch.int32, device=_lib.device())
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
                ok = int(flag.item())
            if not ok:
                self.exchange = exchange = "a2a"
        if exchange != "peer":
            self._send = backend.empty(max(p.local_size, p.local_size_t))
            self._recv = backend.empty(max(p.local_size, p.local_size_t))
            self._xt = backend.empty(p.local_size_t)
            self._yt = backend.empty(p.local_size_t)
        self._pinned_out = None

    def _init_peer(self):
        """XT / YT slabs in symmetric memory + the peer pointer tables of the transpose kernels."""
        import ctypes
        import os

        import torch.distributed._symmetric_memory as symm

        torch = _lib.require_cuda()
        p = self.plan
        group = self.group if self.group is not None else self.dist.group.WORLD
        # symmetric allocations have the same size on every rank: the largest dn-major slab
        n_t = max((p.cols(q)[1] - p.cols(q)[0]) for q in range(self.world)) * p.num_up
        dev = _lib.device()
        self._xt_sym = symm.empty(max(n_t, 1), dtype=torch.float64, device=dev)
        self._yt_sym = symm.empty(max(n_t, 1), dtype=torch.float64, device=dev)
        self._h_xt = symm.rendezvous(self._xt_sym, group)
        self._h_yt = symm.rendezvous(self._yt_sym, group)
        self._xt = self._xt_sym[: max(p.local_size_t, 1)]
        self._yt = self._yt_sym[: max(p.local_size_t, 1)]
        self._peer_xt = (ctypes.c_void_p * self.world)(*[int(a) for a in self._h_xt.buffer_ptrs])
        self._peer_yt = (ctypes.c_void_p * self.world)(*[int(a) for a in self._h_yt.buffer_ptrs])
        self._cb = (ctypes.c_int64 * (self.world + 1))(*p.col_bounds)
        self._side = torch.cuda.Stream()
        # SMs left to the push transpose while the local dn pass runs (the class-major kernel takes
        # one CTA and all shared memory per SM, so without this the two kernels serialise)
        import os
        self._sm_count = torch.cuda.get_device_properties(_lib.device()).multi_processor_count
        self._push_sms = 32   # measured on 2 x B200
        # the same choreography as ONE C call (cmpy_hv_apply_sharded): control block in symmetric memory
        # for the library's own barrier / all-reduce kernels
        self._cdist = None
        if os.environ.get("CMPY_DIST_PYTHON", "") != "1":
            L = _lib.lib()
            nctl = max(int(L.cmpy_dist_ctl_bytes()) // 8, 1)
            self._ctl_sym = symm.empty(nctl, dtype=torch.float64, device=dev)
            self._ctl_sym.zero_()
            self._h_ctl = symm.rendezvous(self._ctl_sym, group)
            torch.cuda.synchronize()
            self._h_ctl.barrier(channel=0)     # every control block is zero before the first handshake
            peer_ctl = (ctypes.c_void_p * self.world)(*[int(a) for a in self._h_ctl.buffer_ptrs])
            handle = ctypes.c_void_p()
            _lib.check(L.cmpy_dist_create(self.backend.op_main.handle, self.backend.op_t.handle, self.world,
                                          self.rank, self._peer_xt, self._peer_yt, peer_ctl, ctypes.byref(handle)),
                       "cmpy_dist_create")
            self._cdist = handle

    def __del__(self):
        h, self._cdist = getattr(self, "_cdist", None), None
        if h:
            try:
                _lib.lib().cmpy_dist_destroy(h)
            except Exception:
                pass

    def _apply_local_peer(self, x_local, out, accumulate=False):
        torch = _lib.require_cuda()
        if self._cdist is not None:
            _lib.check(_lib.lib().cmpy_hv_apply_sharded(self._cdist, _lib.ptr(x_local), _lib.ptr(out),
                                                        int(bool(accumulate)), _lib.stream_ptr()),
                       "cmpy_hv_apply_sharded")
            return out
        p, be, L = self.plan, self.backend, _lib.lib()
        r0, _ = p.rows()
        c0, _ = p.cols()
        nrows, ncols, nu, nd = p.nrows, p.ncols, p.num_up, p.num_dn
        main = torch.cuda.current_stream()
        # every rank is done with the XT / YT slabs of the previous call
        self._h_xt.barrier(channel=0)
        # push the transposed tiles into the owners' XT slabs (side stream) under the local
        # diagonal + dn-hop pass (main stream)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            _lib.check(L.cmpy_transpose_push(_lib.ptr(x_local), nrows, nd, r0, nu, self.world, self._cb,
                                             self._peer_xt, _lib.stream_ptr()), "cmpy_transpose_push")
        limit = self._sm_count - self._push_sms if (self.world > 1 and 0 < self._push_sms < self._sm_count) else 0
        _lib.check(L.cmpy_hubbard_set_grid_limit(be.op_main.handle, limit), "cmpy_hubbard_set_grid_limit")
        if accumulate:
            be.apply_rows(x_local, r0, nrows, out, accumulate=True)
        else:
            be.apply_rows(x_local, r0, nrows, out)
        _lib.check(L.cmpy_hubbard_set_grid_limit(be.op_main.handle, 0), "cmpy_hubbard_set_grid_limit")
        main.wait_stream(self._side)
        return self._apply_second_half(out, r0, c0, nrows, ncols, nu, nd)

    def _apply_second_half(self, out, r0, c0, nrows, ncols, nu, nd):
        be, L = self.backend, _lib.lib()
        self._h_xt.barrier(channel=0)          # all pushes have landed
        be.apply_rows_t(self._xt, c0, ncols, self._yt)   # up hops, row-local in the dn-major slab
        self._h_yt.barrier(channel=0)          # every YT slab is complete
        _lib.check(L.cmpy_transpose_pull_acc(_lib.ptr(out), nrows, nd, r0, nu, self.world, self._cb,
                                             self._peer_yt, _lib.stream_ptr()), "cmpy_transpose_pull_acc")
        return out

    @property
    def local_size(self):
        return self.plan.local_size

    def _all_to_all(self, recv, send, recv_counts, send_counts):
        if self.world == 1:
            recv[: send_counts[0]].copy_(send[: send_counts[0]])
            return
        self.dist.all_to_all_single(recv[: sum(recv_counts)], send[: sum(send_counts)],
                                    output_split_sizes=recv_counts, input_split_sizes=send_counts,
                                    group=self.group)

    def profile_phases(self, x_local, out, reps=5):
        """Per-phase device times (ms, CUDA events, phases serialised) of the peer-memory H.v:
        dn pass, push transpose, up pass, pull transpose, barrier.  Diagnostic only."""
        torch = _lib.require_cuda()
        assert self.exchange == "peer"
        p, be, L = self.plan, self.backend, _lib.lib()
        r0, _ = p.rows(); c0, _ = p.cols()
        nrows, ncols, nu, nd = p.nrows, p.ncols, p.num_up, p.num_dn

        def timed(fn):
            self._h_xt.barrier(channel=0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        res = {}
        res["dn_pass"] = timed(lambda: be.apply_rows(x_local, r0, nrows, out))
        res["push"] = timed(lambda: _lib.check(L.cmpy_transpose_push(
            _lib.ptr(x_local), nrows, nd, r0, nu, self.world, self._cb, self._peer_xt, _lib.stream_ptr())))
        res["up_pass"] = timed(lambda: be.apply_rows_t(self._xt, c0, ncols, self._yt))
        res["pull"] = timed(lambda: _lib.check(L.cmpy_transpose_pull_acc(
            _lib.ptr(out), nrows, nd, r0, nu, self.world, self._cb, self._peer_yt, _lib.stream_ptr())))
        res["barrier"] = timed(lambda: self._h_xt.barrier(channel=0))
        own = nrows * ncols
        res["nvlink_bytes_per_transpose"] = 8 * (p.local_size - own)
        return res

    def apply_local(self, x_local, out=None, accumulate=False):
        """``out = (H x)_local`` (``accumulate``: ``out += (H x)_local``, used by the two-vector
        Lanczos recurrence)."""
        p, be = self.plan, self.backend
        r0, _ = p.rows()
        c0, _ = p.cols()
        nrows, ncols, nu, nd = p.nrows, p.ncols, p.num_up, p.num_dn
        if out is None:
            out = be.empty(p.local_size)
        if self.exchange == "peer":
            return self._apply_local_peer(x_local, out, accumulate)
        # 1. local phase: diagonal + dn hops
        if accumulate:
            be.apply_rows(x_local, r0, nrows, out, accumulate=True)
        else:
            be.apply_rows(x_local, r0, nrows, out)
        # 2. pack transposed blocks, one per peer
        send_counts, recv_counts = p.fwd_send_counts(), p.fwd_recv_counts()
        off = 0
        for q in range(self.world):
            q0, q1 = p.cols(q)
            be.transpose(x_local, q0, nrows, q1 - q0, nd, self._send, off, nrows)
            off += send_counts[q]
        # 3. exchange
        self._all_to_all(self._recv, self._send, recv_counts, send_counts)
        # 4. place the blocks into the dn-major slab XT (ncols x num_up)
        off = 0
        for src in range(self.world):
            s0, s1 = p.rows(src)
            be.copy2d(self._recv, off, ncols, s1 - s0, s1 - s0, self._xt, s0, nu, False)
            off += recv_counts[src]
        # 5. up hops, row-local in the transposed layout
        be.apply_rows_t(self._xt, c0, ncols, self._yt)
        # 6. way back: pack (transpose), exchange, accumulate into y
        off = 0
        for dst in range(self.world):
            d0, d1 = p.rows(dst)
            be.transpose(self._yt, d0, ncols, d1 - d0, nu, self._send, off, ncols)
            off += recv_counts[dst]
        self._all_to_all(self._recv, self._send, send_counts, recv_counts)
        off = 0
        for q in range(self.world):
            q0, q1 = p.cols(q)
            be.copy2d(self._recv, off, nrows, q1 - q0, q1 - q0, out, q0, nd, True)
            off += send_counts[q]
        return out

    def matvec(self, x_local, out=None):
        """Public call on the local slab: CUDA tensor in -> CUDA tensor out; CPU (pinned)
        tensor in -> CPU tensor out (H2D + H.v + D2H).  The host result is a fresh tensor unless
        ``out`` (CPU float64, pinned for full speed) is given."""
        torch = _lib.require_cuda()
        if x_local.is_cuda:
            return self.apply_local(x_local)
        dev = x_local.to(_lib.device(), non_blocking=True)
        y = self.apply_local(dev)
        if out is not None:
            out.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        if self._pinned_out is None:
            self._pinned_out = torch.empty(self.local_size, dtype=torch.float64).pin_memory()
        self._pinned_out.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pinned_out.clone()


def _sharded_matvec_batch(self, xs, outs=None):
    """Pipelined host batch on the local slabs (H2D of slab i+1 || sharded H.v of slab i || D2H of result
    i-1), the sharded counterpart of ``HamiltonOperator.matvec_batch``; collective: every rank passes the
    same number of slabs."""
    from .operators import pipelined_host_batch

    return pipelined_host_batch(self, self.local_size, lambda dx, dy: self.apply_local(dx, out=dy), xs, outs)


ShardedHubbardOperator.matvec_batch = _sharded_matvec_batch


def lanczos_sharded(op, v0_local=None, maxit=500, tol=1e-10, check_every=10, seed=0, callback=None,
                    want_vector=False, resid_tol=0.0):
    """Two-vector Lanczos on an up-string-sharded operator: every rank holds its slab of the two
    Lanczos vectors (plus the operator's XT / YT slabs: 4 slabs in total, which is what lets the
    20-site half-filled sector, 34.1 GB per slab on 8 GPUs, fit 180 GB of HBM).  alpha / beta are
    all-reduced device scalars; the host only looks at them every ``check_every`` iterations.

    Recurrence (same as the single-GPU kernel, ref cmpy/exactdiag.py:324-347 for the
    coefficients): w <- H v - beta_j w (w holds v_{j-1}); alpha_j = <v, w>; w -= alpha_j v;
    beta_{j+1} = |w|; w /= beta_{j+1}; swap.  Returns ``(e0, alpha, beta, nit, converged)``;
    e0 = lowest Ritz value, converged when it moves by less than ``tol`` between two checks (and,
    with ``resid_tol > 0``, the Ritz residual estimate beta_m |s_m| is below ``resid_tol``).
    ``want_vector``: a second pass of the same recurrence accumulates the Ritz vector (one more
    slab); the return value then ends with the local slab of the normalised ground state."""
    import torch
    from scipy.linalg import eigvalsh_tridiagonal

    if (getattr(op, "_cdist", None) is not None and callback is None and not want_vector and resid_tol <= 0):
        res = _lanczos_sharded_c(op, v0_local, maxit, tol, check_every, seed)
        op.last_lanczos_path = "c (cmpy_lanczos_sharded)" if res is not None else "python"
        if res is not None:
            return res
    else:
        op.last_lanczos_path = "python"
    dist = op.dist
    multi = dist.is_initialized() and op.world > 1

    def allsum(t):
        if multi:
            dist.all_reduce(t, group=op.group)
        return t

    def dot(a, b):  # cuBLAS dot is limited to 2^31-1 elements; a slab of the 20-site sector has 4.3e9
        step = 1 << 30
        if a.numel() <= step:
            return torch.dot(a, b)
        acc = torch.zeros((), dtype=a.dtype, device=a.device)
        for i in range(0, a.numel(), step):
            acc += torch.dot(a[i:i + step], b[i:i + step])
        return acc

    n = op.local_size

    def start():
        if v0_local is None:
            dev = op.backend.empty(1).device
            g = torch.Generator(device=dev)
            g.manual_seed(int(seed) + 7919 * op.rank)
            v_ = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        else:
            v_ = v0_local.clone()
        v_.div_(torch.sqrt(allsum(dot(v_, v_))))
        return v_

    v = start()
    w = torch.zeros_like(v)
    alphas = torch.zeros(int(maxit), dtype=torch.float64, device=v.device)
    betas = torch.zeros(int(maxit), dtype=torch.float64, device=v.device)
    e_prev, e0, converged, nit = None, float("nan"), False, 0
    for j in range(int(maxit)):
        if j > 0:
            w.mul_(-betas[j - 1])
        op.apply_local(v, out=w, accumulate=True)
        a = allsum(dot(v, w))
        w.addcmul_(v, a, value=-1.0)
        b = torch.sqrt(allsum(dot(w, w)))
        alphas[j] = a
        betas[j] = b
        nit = j + 1
        last = nit == int(maxit)
        if nit % int(check_every) == 0 or last:
            ah, bh = alphas[:nit].cpu().numpy(), betas[:nit].cpu().numpy()
            resid = 0.0
            if nit > 1:
                if resid_tol > 0:
                    from scipy.linalg import eigh_tridiagonal as _eight

                    ev_, sv_ = _eight(ah, bh[:nit - 1], select="i", select_range=(0, 0))
                    e0, resid = float(ev_[0]), float(bh[nit - 1] * abs(sv_[-1, 0]))
                else:
                    e0 = float(eigvalsh_tridiagonal(ah, bh[:nit - 1], select="i", select_range=(0, 0))[0])
            else:
                e0 = float(ah[0])
            if callback is not None:
                callback(nit, e0)
            if bh[nit - 1] < 1e-14 * max(1.0, abs(e0)):   # invariant subspace: exact
                converged = True
                break
            if e_prev is not None and abs(e0 - e_prev) < tol and (resid_tol <= 0 or resid <= resid_tol):
                converged = True
                break
            e_prev = e0
        w.div_(b)
        v, w = w, v
    ah, bh = alphas[:nit].cpu().numpy(), betas[:nit].cpu().numpy()
    # usable length: a Krylov breakdown (invariant subspace) between two checks leaves a beta ~ 0 followed by
    # meaningless coefficients -- cut there, like cmpy_lanczos_sharded and the single-GPU driver do
    m_use, scale = nit, 0.0
    for i in range(nit):
        if not (np.isfinite(ah[i]) and np.isfinite(bh[i])):
            m_use = i
            break
        scale = max(scale, abs(float(ah[i])) + abs(float(bh[i])))
        if bh[i] <= 1e-13 * max(scale, 1.0):
            m_use = i + 1
            break
    if m_use < nit:
        nit = max(m_use, 1)
        ah, bh = ah[:nit].copy(), bh[:nit].copy()
        e0 = (float(eigvalsh_tridiagonal(ah, bh[:nit - 1], select="i", select_range=(0, 0))[0])
              if nit > 1 else float(ah[0]))
        converged = True
    if not want_vector:
        return e0, ah, bh, nit, converged
    # second pass: psi = sum_j s_j v_j with s the lowest eigenvector of the Lanczos matrix
    from scipy.linalg import eigh_tridiagonal

    if nit > 1:
        _, svec = eigh_tridiagonal(ah, bh[:nit - 1], select="i", select_range=(0, 0))
        coef = svec[:, 0]
    else:
        coef = np.ones(1)
    del v, w
    v = start()
    w = torch.zeros_like(v)
    psi = torch.zeros_like(v)
    for j in range(nit):
        psi.add_(v, alpha=float(coef[j]))
        if j + 1 == nit:
            break
        if j > 0:
            w.mul_(-float(bh[j - 1]))
        op.apply_local(v, out=w, accumulate=True)
        w.add_(v, alpha=-float(ah[j]))
        w.div_(float(bh[j]))
        v, w = w, v
    psi.div_(torch.sqrt(allsum(dot(psi, psi))))
    return e0, ah, bh, nit, converged, psi


def _lanczos_sharded_c(op, v0_local, maxit, tol, check_every, seed):
    """The recurrence of ``lanczos_sharded`` as one C call (``cmpy_lanczos_sharded``: fused device kernels,
    device-side all-reduces over the control blocks).  ``None`` when the operator has no row engine."""
    import ctypes

    torch = _lib.require_cuda()
    n = op.local_size
    dev = _lib.device()
    if v0_local is None:
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed) + 7919 * op.rank)
        r = torch.randn(max(n, 1), dtype=torch.float64, device=dev, generator=g)
    else:
        r = v0_local.clone()
    w = torch.empty_like(r)
    cap = int(maxit) + int(check_every) + 4
    alpha = np.zeros(cap, dtype=np.float64)
    beta = np.zeros(cap + 1, dtype=np.float64)
    nit, e0 = ctypes.c_int(0), ctypes.c_double(0.0)
    rc = _lib.lib().cmpy_lanczos_sharded(
        op._cdist, _lib.ptr(r), _lib.ptr(w), int(maxit), float(tol), int(check_every),
        alpha.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), beta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        ctypes.byref(nit), ctypes.byref(e0), _lib.stream_ptr())
    if rc == _lib.CMPY_ERR_UNSUPPORTED:
        return None
    if rc not in (_lib.CMPY_OK, _lib.CMPY_ERR_NOT_CONVERGED):
        _lib.check(rc, "cmpy_lanczos_sharded")
    m = nit.value
    return e0.value, alpha[:m].copy(), beta[1:m + 1].copy(), m, rc == _lib.CMPY_OK


def gf_continued_fraction_sharded(model, z, pos=0, n_up=None, n_dn=None, num_coeffs=600, tol=1e-12,
                                  group=None, return_info=False):
    """Zero-temperature G_{pos,DN}(z) on an up-string-sharded sector: sharded Lanczos ground state,
    c^+_{pos,dn} / c_{pos,dn} applied slab-locally (a dn ladder operator does not move amplitudes
    between up-rows), sharded Lanczos from the two start vectors, continued fraction on every rank.
    Same convention as ``exactdiag.gf_continued_fraction(..., sigma=DN)`` (signless operators of the
    reference, cmpy/operators.py:652-703; for n_up = n_dn it equals the sigma=UP function)."""
    from .basis import DN, Sector
    from .exactdiag import _z_tensor, cf_eval
    from .operators import AnnihilationOperator, CreationOperator

    torch = _lib.require_cuda()
    basis = model.basis
    L = basis.num_sites
    n_up = L // 2 if n_up is None else n_up
    n_dn = L // 2 if n_dn is None else n_dn
    op = ShardedHubbardOperator(model, n_up, n_dn, group=group)
    e0, _, _, nit0, conv, psi = lanczos_sharded(op, maxit=3000, tol=tol, check_every=10, want_vector=True,
                                                resid_tol=1e-10)
    r0, r1 = op.plan.rows()
    full = basis.get_sector(n_up, n_dn)
    up_slab = np.asarray(full.up_states)[r0:r1]
    sec_slab = Sector(up_slab, np.asarray(full.dn_states), n_up, n_dn, L)
    zt, zshape = _z_tensor(z)
    g = torch.zeros_like(zt)
    info = {"e0": e0, "gs_iterations": nit0, "gs_converged": bool(conv), "norms": [0.0, 0.0], "nit": [0, 0]}
    del op
    for part, (nd_t, sign, cls) in enumerate(((n_dn + 1, +1, CreationOperator), (n_dn - 1, -1, AnnihilationOperator))):
        if nd_t < 0 or nd_t > L:
            continue
        sec_t = Sector(up_slab, np.asarray(basis.get_states(nd_t)), n_up, nd_t, L)
        phi = cls(sec_slab, sec_t, pos=pos, sigma=DN).apply(psi)
        nrm = torch.zeros((), dtype=phi.dtype, device=phi.device)
        for i0 in range(0, phi.numel(), 1 << 30):   # cuBLAS dot is limited to 2^31 - 1 elements
            nrm += torch.dot(phi[i0:i0 + (1 << 30)], phi[i0:i0 + (1 << 30)])
        if op_world(group) > 1:
            import torch.distributed as dist

            dist.all_reduce(nrm, group=group)
        norm2 = float(nrm)
        info["norms"][part] = norm2
        if norm2 < 1e-28:
            continue
        op_t = ShardedHubbardOperator(model, n_up, nd_t, group=group)
        m = min(int(num_coeffs), op_t.shape[0])
        _, al, be, nit, _ = lanczos_sharded(op_t, v0_local=phi, maxit=m, tol=0.0, check_every=m)
        info["nit"][part] = nit
        cf_eval(al, be[:nit - 1], norm2, e0, zt, sign=sign, out=g)
        del op_t
    out = g.cpu().numpy().reshape(zshape)
    return (out, info) if return_info else out


def op_world(group=None):
    import torch.distributed as dist

    return dist.get_world_size(group) if dist.is_initialized() else 1
