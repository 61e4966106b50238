# -*- coding: utf-8 -*-
"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (numpy restatement of cmpy's hot path).

This module restates, in plain numpy, the algorithms of the reference
(dylanljones/cmpy, `/root/reference`) that the CUDA engine in `cmpy_b200/`
replaces.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline legs may import it.  Nothing under `cmpy_b200/`
imports or calls it, and the product path has no CPU fallback.

Parity status: PINNED.  Every function here is checked against
  (a) the reference's own golden vectors / known-answer tests
      (cmpy/tests/test_models_hubbard.py:14-25, test_basis.py:124-141,
       test_operator.py:15-28, test_models_heisenberg.py:15-35), and
  (b) outputs of the unmodified reference run in the dev container, committed
      as fixtures under `tests/golden/` by `oracle/make_golden.py`
(see tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py).

Each function cites the reference file:line it follows.
"""
import numpy as np

UP, DN = 1, 2  # cmpy/basis.py:40


# ---------------------------------------------------------------------------
# Sector enumeration / ranking
# ---------------------------------------------------------------------------

def binom_table(nmax=64):
    """Pascal triangle C[n][k] as python ints (exact)."""
    c = [[0] * (nmax + 1) for _ in range(nmax + 1)]
    for n in range(nmax + 1):
        c[n][0] = 1
        for k in range(1, n + 1):
            c[n][k] = c[n - 1][k - 1] + c[n - 1][k]
    return c


_BINOM = binom_table(64)


def enumerate_states(num_sites, n):
    """All `num_sites`-bit integers with popcount `n`, ascending (int64 ndarray).

    Restates `Basis.generate_states` (cmpy/basis.py:655-666) -- the reference builds
    the set from `itertools.permutations` and sorts it; the result is the ascending
    list of fixed-popcount integers, produced here with Gosper's hack.
    (The reference returns python lists for n in (None, 0, 1); the *values* agree.)
    """
    if n is None:
        return np.arange(2 ** num_sites, dtype=np.int64)
    if n < 0 or n > num_sites:
        return np.zeros(0, dtype=np.int64)
    count = _BINOM[num_sites][n]
    out = np.empty(count, dtype=np.int64)
    if n == 0:
        out[0] = 0
        return out
    s = (1 << n) - 1
    for i in range(count):
        out[i] = s
        c = s & -s
        r = s + c
        s = (((r ^ s) >> 2) // c) | r
    return out


def rank_state(state, n=None):
    """Colex (combinadic) rank of `state` among integers of equal popcount.

    Equals `bisect_left(states, state)` (cmpy/operators.py:276-299) on the ascending
    fixed-popcount list.
    """
    r, k, p = 0, 0, 0
    s = int(state)
    while s:
        if s & 1:
            k += 1
            r += _BINOM[p][k]
        s >>= 1
        p += 1
    return r


def rank_states(states_sorted, queries):
    """Vectorised `bisect_left` (cmpy/operators.py:276-299)."""
    return np.searchsorted(np.asarray(states_sorted), np.asarray(queries), side="left")


def spin_states(num_sites, s):
    """`SpinBasis.generate_states(s)` (cmpy/basis.py:748-764): ascending ints with
    popcount num_sites/2 + s; ValueError when unrealisable."""
    if s is None:
        return list(range(2 ** num_sites))
    n_up = num_sites / 2 + s
    n_dn = num_sites / 2 - s
    if (n_up % 1 != 0.0) or (n_dn % 1 != 0.0):
        raise ValueError(f"Total spin of {s} not realizable with {num_sites} sites")
    return [int(v) for v in enumerate_states(num_sites, int(n_up))]


# ---------------------------------------------------------------------------
# Matrix elements (projectors), exact emission order of the reference
# ---------------------------------------------------------------------------

def weighted_element(state, values):
    """cmpy/operators.py:226-250 -- sum of values[i] over set bits, ascending i."""
    value = 0.0
    for i in range(len(values)):
        if int(state) & (1 << i):
            value += float(values[i])
    return value


def weighted_elements(states, values):
    """Vectorised `weighted_element` with the same (ascending-site) summation order."""
    states = np.asarray(states, dtype=np.int64)
    out = np.zeros(states.shape, dtype=np.float64)
    for i in range(len(values)):
        bit = ((states >> i) & 1).astype(bool)
        out = np.where(bit, out + float(values[i]), out)
    return out


def onsite_triplets(up_states, dn_states, eps):
    """`project_onsite_energy` (cmpy/operators.py:359-422): up block then dn block,
    zero energies skipped. Returns (rows, cols, vals)."""
    up_states = np.asarray(up_states, dtype=np.int64)
    dn_states = np.asarray(dn_states, dtype=np.int64)
    num_up, num_dn = len(up_states), len(dn_states)
    e_up = weighted_elements(up_states, eps)
    e_dn = weighted_elements(dn_states, eps)
    all_dn = np.arange(num_dn, dtype=np.int64)
    all_up = np.arange(num_up, dtype=np.int64)
    sel_up = np.nonzero(e_up != 0.0)[0]
    idx_up = (sel_up[:, None] * num_dn + all_dn[None, :]).ravel()
    val_up = np.repeat(e_up[sel_up], num_dn)
    sel_dn = np.nonzero(e_dn != 0.0)[0]
    idx_dn = (all_up[None, :] * num_dn + sel_dn[:, None]).ravel()
    val_dn = np.repeat(e_dn[sel_dn], num_up)
    idx = np.concatenate([idx_up, idx_dn])
    return idx, idx.copy(), np.concatenate([val_up, val_dn])


def inter_triplets(up_states, dn_states, u):
    """`project_hubbard_inter` (cmpy/operators.py:305-356): up-major, zeros skipped."""
    up_states = np.asarray(up_states, dtype=np.int64)
    dn_states = np.asarray(dn_states, dtype=np.int64)
    both = up_states[:, None] & dn_states[None, :]
    energy = weighted_elements(both, u).ravel()
    idx = np.nonzero(energy != 0.0)[0].astype(np.int64)
    return idx, idx.copy(), energy[idx]


def hopping_sign(state, width, site1, site2):
    """`_hopping_sign` + `bit_count(number, width)` (cmpy/operators.py:253-273,425-433):
    (-1)^popcount(state & mask(site1+1..site2-1) & (2^width-1))."""
    mask = 0
    for i in range(site1 + 1, site2):
        mask += 1 << i
    masked = int(state) & mask
    count = 0
    for i in range(width):
        if masked & (1 << i):
            count += 1
    return (-1) ** count


def species_hops(states, width, site1, site2):
    """`_compute_hopping_term` (cmpy/operators.py:436-460) without the value:
    returns (origin_idx, target_idx, sign) for every state whose bits at site1,
    site2 differ, ascending origin."""
    assert site1 < site2
    states = np.asarray(states, dtype=np.int64)
    b1 = (states >> site1) & 1
    b2 = (states >> site2) & 1
    sel = np.nonzero(b1 != b2)[0]
    new = states[sel] ^ ((1 << site1) | (1 << site2))
    tgt = rank_states(states, new)
    between = 0
    for i in range(site1 + 1, site2):
        if i < width:
            between |= 1 << i
    cnt = np.zeros(len(sel), dtype=np.int64)
    masked = states[sel] & between
    for i in range(max(site2, 1)):
        cnt += (masked >> i) & 1
    sign = 1 - 2 * (cnt & 1)
    return sel.astype(np.int64), tgt.astype(np.int64), sign.astype(np.int64)


def hopping_triplets(up_states, dn_states, num_sites, site1, site2, hop):
    """`project_hopping` (cmpy/operators.py:463-527): up block (each hop expanded over
    all dn) followed by the dn block (each hop expanded over all up)."""
    num_up, num_dn = len(up_states), len(dn_states)
    all_dn = np.arange(num_dn, dtype=np.int64)
    all_up = np.arange(num_up, dtype=np.int64)
    o, t, s = species_hops(up_states, num_sites, site1, site2)
    rows_u = (o[:, None] * num_dn + all_dn[None, :]).ravel()
    cols_u = (t[:, None] * num_dn + all_dn[None, :]).ravel()
    vals_u = np.repeat(s * hop, num_dn).astype(np.float64)
    o, t, s = species_hops(dn_states, num_sites, site1, site2)
    rows_d = (all_up[None, :] * num_dn + o[:, None]).ravel()
    cols_d = (all_up[None, :] * num_dn + t[:, None]).ravel()
    vals_d = np.repeat(s * hop, num_up).astype(np.float64)
    return (np.concatenate([rows_u, rows_d]), np.concatenate([cols_u, cols_d]),
            np.concatenate([vals_u, vals_d]))


def _cat(parts):
    rows = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.int64)
    cols = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int64)
    vals = np.concatenate([p[2] for p in parts]) if parts else np.zeros(0, np.float64)
    return rows.astype(np.int64), cols.astype(np.int64), vals.astype(np.float64)


def hubbard_triplets(up_states, dn_states, num_sites, neighbors, inter, eps, hop):
    """`_ham_data` (cmpy/models/hubbard.py:13-22): onsite, interaction, then one
    `project_hopping` per neighbor pair with i<j. `eps` is already eps-mu
    (hubbard.py:77)."""
    energy = np.full(num_sites, eps, dtype=np.float64)
    interaction = np.full(num_sites, inter, dtype=np.float64)
    parts = [onsite_triplets(up_states, dn_states, energy),
             inter_triplets(up_states, dn_states, interaction)]
    for i, j in neighbors:
        if i < j:
            parts.append(hopping_triplets(up_states, dn_states, num_sites, i, j, hop))
    return _cat(parts)


def siam_triplets(up_states, dn_states, u, eps_imp, eps_bath, v, mu=None):
    """`SingleImpurityAndersonModel._hamiltonian_data` (cmpy/models/anderson.py:147-158):
    sign width 0 (signless hops), u=[U,0..], eps=[eps_imp-mu, eps_bath..]; mu=u/2 when
    None (anderson.py:61)."""
    mu = u / 2 if mu is None else mu
    eps_bath = np.atleast_1d(eps_bath).astype(np.float64)
    v = np.atleast_1d(v).astype(np.float64)
    if len(eps_bath) > 1 and len(v) == 1:
        v = np.ones(len(eps_bath)) * v[0]
    if len(eps_bath) == 1 and len(v) > 1:
        eps_bath = np.ones(len(v)) * eps_bath[0]
    num_bath = len(eps_bath)
    uu = np.append(u, np.zeros(num_bath))
    eps = np.append(eps_imp - mu, eps_bath)
    parts = [onsite_triplets(up_states, dn_states, eps),
             inter_triplets(up_states, dn_states, uu)]
    for j in range(num_bath):
        parts.append(hopping_triplets(up_states, dn_states, 0, 0, j + 1, v[j]))
    return _cat(parts)


def heisenberg_triplets(states, neighbor_lists, j=1.0, jz=1.0):
    """`HeisenbergModel._hamiltonian_data` (cmpy/models/heisenberg.py:19-40): per
    state, per directed neighbor pair: diagonal sign*0.25*jz, and for anti-parallel
    bits an off-diagonal 0.25*j/2 to the flipped state. Emission order preserved."""
    states = [int(s) for s in states]
    srt = np.asarray(states, dtype=np.int64)
    is_sorted = bool(np.all(srt[:-1] < srt[1:])) if len(srt) > 1 else True
    index = {s: i for i, s in enumerate(states)}
    factor = 0.25
    rows, cols, vals = [], [], []
    num_sites = len(neighbor_lists)
    for idx1, s1 in enumerate(states):
        for pos1 in range(num_sites):
            for pos2 in neighbor_lists[pos1]:
                b1 = (s1 >> pos1) & 1
                b2 = (s1 >> pos2) & 1
                sign = (-1) ** b1 * (-1) ** b2
                rows.append(idx1); cols.append(idx1); vals.append(sign * factor * jz)
                if b1 != b2:
                    s2 = s1 ^ (1 << pos1) ^ (1 << pos2)
                    rows.append(idx1); cols.append(index[s2]); vals.append(factor * j / 2)
    del is_sorted
    return (np.asarray(rows, np.int64), np.asarray(cols, np.int64),
            np.asarray(vals, np.float64))


# ---------------------------------------------------------------------------
# H.v
# ---------------------------------------------------------------------------

def coo_matvec(size, rows, cols, vals, x):
    """`HamiltonOperator._matvec` (cmpy/operators.py:626-630):
    y[col] += val * x[row], duplicates accumulate."""
    y = np.zeros(size, dtype=np.result_type(vals.dtype, x.dtype))
    np.add.at(y, cols, vals * x[rows])
    return y


def coo_dense(size, rows, cols, vals):
    """`HamiltonOperator.toarray` (cmpy/operators.py:632-635): duplicates summed."""
    a = np.zeros((size, size), dtype=np.float64)
    np.add.at(a, (rows, cols), vals)
    return a


def hubbard_matvec_free(up_states, dn_states, neighbors, inter, eps, hop, x, width=None,
                        u_sites=None, eps_sites=None, hop_bonds=None):
    """Matrix-free H.x for the Hubbard / SIAM sector, numerically the same operator
    as `hubbard_triplets`+`coo_matvec` but O(dim) memory (used for sizes where the
    triplet list is too large). Follows cmpy/operators.py:305-527, models/hubbard.py:13-22.

    `width` = sign width (`num_sites` argument of project_hopping; SIAM passes 0).
    """
    up_states = np.asarray(up_states, dtype=np.int64)
    dn_states = np.asarray(dn_states, dtype=np.int64)
    num_up, num_dn = len(up_states), len(dn_states)
    nsites = int(max(up_states.max(initial=0), dn_states.max(initial=0))).bit_length()
    bonds = [(int(i), int(j)) for i, j in neighbors if i < j]
    for i, j in bonds:
        nsites = max(nsites, j + 1)
    if width is None:
        width = nsites
    eps_arr = np.full(nsites, eps, dtype=np.float64) if eps_sites is None else np.asarray(eps_sites, float)
    u_arr = np.full(nsites, inter, dtype=np.float64) if u_sites is None else np.asarray(u_sites, float)
    hops = [hop] * len(bonds) if hop_bonds is None else list(hop_bonds)
    X = np.asarray(x).reshape(num_up, num_dn)
    e_up = weighted_elements(up_states, eps_arr)
    e_dn = weighted_elements(dn_states, eps_arr)
    Y = (e_up[:, None] + e_dn[None, :]) * X
    # interaction, row by row to bound memory
    for a in range(num_up):
        Y[a] += weighted_elements(up_states[a] & dn_states, u_arr) * X[a]
    for (i, j), t in zip(bonds, hops):
        o, tg, s = species_hops(up_states, width, i, j)
        # y[col] += val * x[row]  with row = origin, col = target
        np.add.at(Y, tg, (s * t)[:, None] * X[o])
        o, tg, s = species_hops(dn_states, width, i, j)
        YT = Y.T
        np.add.at(YT, tg, (s * t)[:, None] * X.T[o])
    return Y.reshape(-1)


# ---------------------------------------------------------------------------
# Ladder operators (signless, as in the reference)
# ---------------------------------------------------------------------------

def ladder_apply(x, up_states, dn_states, up_states_t, dn_states_t, pos, sigma, dagger):
    """`_apply_creation_up/dn`, `_apply_annihilation_up/dn` (cmpy/operators.py:652-703):
    y[rank'(s^bit), d] = x[s, d] when the bit may be created/annihilated. NO fermionic
    sign. For sigma=DN the target row stride is the TARGET sector's num_dn (the
    reference uses the origin's, operators.py:668,675 -- an IndexError bug; parity is
    pinned for UP only)."""
    up_states = np.asarray(up_states, np.int64); dn_states = np.asarray(dn_states, np.int64)
    up_t = np.asarray(up_states_t, np.int64); dn_t = np.asarray(dn_states_t, np.int64)
    nu, nd = len(up_states), len(dn_states)
    nut, ndt = len(up_t), len(dn_t)
    X = np.asarray(x).reshape(nu, nd)
    op = 1 << pos
    Y = np.zeros((nut, ndt), dtype=X.dtype)
    if sigma == UP:
        occ = (up_states & op) != 0
        sel = np.nonzero(~occ if dagger else occ)[0]
        tgt = rank_states(up_t, up_states[sel] ^ op)
        Y[tgt, :] = X[sel, :]
    else:
        occ = (dn_states & op) != 0
        sel = np.nonzero(~occ if dagger else occ)[0]
        tgt = rank_states(dn_t, dn_states[sel] ^ op)
        Y[:, tgt] = X[:, sel]
    return Y.reshape(-1)


# ---------------------------------------------------------------------------
# Lanczos / continued fraction / Lehmann
# ---------------------------------------------------------------------------

def lanczos_coeffs_normalised(matvec, v0, m):
    """Plain (no re-orthogonalisation) Lanczos, normalised 3-term recurrence.
    Mathematically the recurrence of `iter_lanczos_coeffs` (cmpy/exactdiag.py:324-347)
    in the normalised basis: returns (alpha[m'], beta[m'-1], norm0)."""
    v = np.asarray(v0, dtype=np.float64)
    n0 = np.linalg.norm(v)
    v = v / n0
    v_prev = np.zeros_like(v)
    alphas, betas = [], []
    beta = 0.0
    for _ in range(m):
        w = matvec(v) - beta * v_prev
        a = float(np.dot(v, w))
        w = w - a * v
        alphas.append(a)
        beta = float(np.linalg.norm(w))
        if beta < 1e-13:
            break
        betas.append(beta)
        v_prev, v = v, w / beta
    return np.asarray(alphas), np.asarray(betas[:len(alphas) - 1]), n0


def reference_lanczos_coeffs(ham, psi0, size):
    """Verbatim arithmetic of `iter_lanczos_coeffs` (cmpy/exactdiag.py:324-347) for a
    dense matrix, with the start vector passed in instead of drawn from the global RNG.
    Returns (a[size], b[size-1])."""
    psi = np.asarray(psi0, dtype=np.float64)
    a_list, b_list = [], []
    a = np.dot(psi, np.dot(ham, psi)) / np.dot(psi, psi)
    a_list.append(a)
    psi_new = np.dot(ham, psi) - a * psi
    psi_prev, psi = psi, psi_new
    for _ in range(1, size):
        a = np.dot(psi, np.dot(ham, psi)) / np.dot(psi, psi)
        b2 = np.dot(psi, psi) / np.dot(psi_prev, psi_prev)
        psi_new = np.dot(ham, psi) - a * psi - b2 * psi_prev
        a_list.append(a)
        b_list.append(np.sqrt(b2))
        psi_prev, psi = psi, psi_new
    return np.asarray(a_list), np.asarray(b_list)


def tridiag_lowest(alpha, beta, k=1):
    import scipy.linalg as la
    k = min(k, len(alpha))
    if len(alpha) == 1:
        return np.asarray(alpha[:1]), np.ones((1, 1))
    return la.eigh_tridiagonal(alpha, beta, select="i", select_range=(0, k - 1))


def cf_eval(alpha, beta, norm2, zshift):
    """norm2 / (zs - a0 - b1^2/(zs - a1 - ...)) evaluated bottom-up; zs complex array."""
    zs = np.asarray(zshift, dtype=np.complex128)
    m = len(alpha)
    g = zs - alpha[m - 1]
    for k in range(m - 2, -1, -1):
        g = zs - alpha[k] - (beta[k] ** 2) / g
    return norm2 / g


def zero_t_lehmann(z, e0, gs, evals_p1, evecs_p1, cdag_gs, evals_m1, evecs_m1, c_gs):
    """T=0 Lehmann sum assembled from reference parts:
    G(z) = sum_m |<m|c^dag gs>|^2/(z-E_m+E0) + sum_n |<n|c gs>|^2/(z+E_n-E0).
    It is the beta->inf limit of `_accumulate_sum` (cmpy/exactdiag.py:110-129)."""
    z = np.asarray(z, np.complex128)
    g = np.zeros_like(z)
    if evals_p1 is not None:
        w = np.abs(evecs_p1.T @ cdag_gs) ** 2
        g += (w[None, :] / (z[:, None] - evals_p1[None, :] + e0)).sum(axis=1)
    if evals_m1 is not None:
        w = np.abs(evecs_m1.T @ c_gs) ** 2
        g += (w[None, :] / (z[:, None] + evals_m1[None, :] - e0)).sum(axis=1)
    return g


def gf0_lehmann(ham, z, mu=0.0):
    """`gf0_lehmann(..., mode='diag')` (cmpy/greens.py:18-64):
    G_ii(z) = sum_k |v_ik|^2 / (z + mu - eps_k); returns (Nz, N)."""
    eigvals, eigvecs = np.linalg.eigh(np.asarray(ham, dtype=np.float64))
    z = np.atleast_1d(z)
    arg = np.subtract.outer(z + mu, eigvals)
    # verbatim contraction of greens.py:51-64 ('diag'): out[..., i] =
    # sum_j adj[i, j] * (1/arg)[..., j] * vecs[j, i] = sum_j |vecs[j, i]|^2 / arg_j
    eigvecs_adj = np.conj(eigvecs).T
    return np.einsum("ij,...j,ji->...i", eigvecs_adj, 1 / arg, eigvecs)


def gf_lehmann_finite_t(sector_solver, num_sites, z, beta, pos=0, sigma=UP):
    """Finite-temperature Lehmann sum, restating `gf_lehmann` +
    `GreensFunctionMeasurement.accumulate` + `_accumulate_sum`
    (cmpy/exactdiag.py:110-129, 197-245) including the running-minimum rescaling.

    `sector_solver(n_up, n_dn)` -> (evals, evecs, up_states, dn_states).
    Returns (gf, part_scaled, gs_energy, occ, occ_double)."""
    z = np.asarray(z, np.complex128)
    gf = np.zeros_like(z)
    part = 0.0
    e_gs = np.inf
    occ = 0.0
    occ2 = 0.0
    cache = {}

    def solve(nu, nd):
        if (nu, nd) not in cache:
            cache[(nu, nd)] = sector_solver(nu, nd)
        return cache[(nu, nd)]

    for n_up in range(num_sites + 1):
        for n_dn in range(num_sites + 1):
            if sigma == UP:
                if n_up >= num_sites:
                    continue
                p1 = (n_up + 1, n_dn)
            else:
                if n_dn >= num_sites:
                    continue
                p1 = (n_up, n_dn + 1)
            evals, evecs, ups, dns = solve(n_up, n_dn)
            evals1, evecs1, ups1, dns1 = solve(*p1)
            emin = evals.min()
            factor = 1.0
            if emin < e_gs:
                factor = np.exp(-beta * (e_gs - emin))
                e_gs = emin
            part = part * factor + np.sum(np.exp(-beta * (evals - e_gs)))
            if factor != 1.0:
                gf = gf * factor
            cdag_evec = np.stack(
                [ladder_apply(evecs[:, k], ups, dns, ups1, dns1, pos, sigma, True)
                 for k in range(evecs.shape[1])], axis=1)
            overlap = np.abs(evecs1.T.conj() @ cdag_evec) ** 2  # (m, n)
            ex = np.exp(-beta * (evals - e_gs))
            ex1 = np.exp(-beta * (evals1 - e_gs))
            for m in range(len(evals1)):
                zm = z - evals1[m]
                wts = overlap[m, :] * (ex + ex1[m])
                gf = gf + (wts[None, :] / (zm[:, None] + evals[None, :])).sum(axis=1)
            # occupations (cmpy/exactdiag.py:60-107)
            nd = len(dns)
            W = (np.abs(evecs) ** 2) @ ex  # weight per basis state
            W = W.reshape(len(ups), nd)
            if sigma == UP:
                m_occ = ((np.asarray(ups) >> pos) & 1).astype(bool)
                o = W[m_occ, :].sum()
            else:
                m_occ = ((np.asarray(dns) >> pos) & 1).astype(bool)
                o = W[:, m_occ].sum()
            both = ((np.asarray(ups)[:, None] & np.asarray(dns)[None, :]) >> pos) & 1
            occ = occ * factor + o
            occ2 = occ2 * factor + (W * both).sum()
    return gf / part, part, e_gs, occ / part, occ2 / part


# ---------------------------------------------------------------------------
# helpers shared by tests / bench
# ---------------------------------------------------------------------------

def gf_realtime_dense(num_sites, neighbors, inter, eps, hop, n_up, n_dn, gs_energy, gs_state, times,
                      pos=0, greater=True):
    """`gf_greater` / `gf_lesser` (cmpy/exactdiag.py:248-301, sigma=UP) with a dense
    eigendecomposition of the target-sector Hamiltonian in place of `expm_multiply`:
    greater: -i e^{+i E0 t} <phi| e^{-i H t} |phi>, phi = c^+_pos |gs> (signless ladder);
    lesser:  +i e^{-i E0 t} <phi| e^{+i H t} |phi>, phi = c_pos |gs>."""
    up, dn = enumerate_states(num_sites, n_up), enumerate_states(num_sites, n_dn)
    up_t = enumerate_states(num_sites, n_up + 1 if greater else n_up - 1)
    phi = ladder_apply(np.asarray(gs_state, dtype=np.float64), up, dn, up_t, dn, pos, 1, bool(greater))
    r, c, v = hubbard_triplets(up_t, dn, num_sites, neighbors, inter, eps, hop)
    ev, vecs = np.linalg.eigh(coo_dense(len(up_t) * len(dn), r, c, v))
    a2 = (vecs.T @ phi) ** 2
    sgn = -1.0 if greater else 1.0
    times = np.asarray(times, dtype=np.float64)
    overlaps = (np.exp(1j * sgn * np.outer(times, ev)) * a2[None, :]).sum(axis=1)
    return (1j * sgn) * np.exp(-1j * sgn * gs_energy * times) * overlaps


def chain_neighbors(num_sites, periodic=False):
    nb = [[i, i + 1] for i in range(num_sites - 1)]
    if periodic and num_sites > 2:
        nb.append([0, num_sites - 1])
    return nb


def square_neighbors(nx, ny):
    """Open nx x ny square lattice, site = nx*row + col."""
    nb = []
    for r in range(ny):
        for c in range(nx):
            i = nx * r + c
            if c + 1 < nx:
                nb.append([i, i + 1])
            if r + 1 < ny:
                nb.append([i, i + nx])
    return nb


def chain_neighbor_lists(num_sites, periodic=False):
    out = []
    for i in range(num_sites):
        if periodic and num_sites > 2:
            out.append([(i - 1) % num_sites, (i + 1) % num_sites])
        else:
            out.append([k for k in (i - 1, i + 1) if 0 <= k < num_sites])
    return out


# ---------------------------------------------------------------------------
# spin observables (reference: scripts/heisenberg.py)
# ---------------------------------------------------------------------------

def sz_expval(states, gs, pos=0):
    """`sz_expval` (scripts/heisenberg.py:50-57): sum of a^2 * (+1/2 if bit `pos` set else -1/2)."""
    sz = 0.0
    for ai, si in zip(gs, states):
        b = (int(si) >> pos) & 1
        sz += (-1) ** (b + 1) / 2 * ai * ai
    return sz


def sz_correl(states, gs, delta, j=1.0, pos=0):
    """`sz_correl` (scripts/heisenberg.py:131-138, there with pos = 0): sum of a^2 * sign * j / 4,
    sign +1 when the bits at `pos` and `pos + delta` are equal."""
    res = 0.0
    for ai, si in zip(gs, states):
        b1 = (int(si) >> pos) & 1
        b2 = (int(si) >> (pos + delta)) & 1
        res += ai * ai * (-1) ** b1 * (-1) ** b2 * j / 4
    return res
