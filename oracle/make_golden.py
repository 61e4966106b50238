# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/reference_golden.npz` by running
the UNMODIFIED reference (dylanljones/cmpy at /root/reference) through `oracle/refshim.py`.

Run in the dev container (the reference is not present on the GPU box):

    NUMBA_DISABLE_JIT=1 python oracle/make_golden.py

`NUMBA_DISABLE_JIT=1` mirrors the reference CI (.github/workflows/test.yml:10-11).
The fixtures are the parity anchors for `oracle/oracle_np.py` (CPU tests) and for the
CUDA engine (GPU tests).
"""
import os
import sys
import time

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import refshim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "reference_golden.npz")


def chain(n, periodic=False):
    nb = [[i, i + 1] for i in range(n - 1)]
    if periodic and n > 2:
        nb.append([0, n - 1])
    return nb


def main():
    t00 = time.time()
    cmpy = refshim.load_reference()
    from cmpy.basis import Basis, UP, DN
    from cmpy import operators as ops
    from cmpy.models import HubbardModel, HeisenbergModel, SingleImpurityAndersonModel
    from cmpy import exactdiag as ed
    from cmpy.greens import gf0_lehmann

    g = {}

    # 1. sector enumeration (basis.py:655-676), ordering + values
    for L in range(1, 10):
        b = Basis(L)
        for n in range(L + 1):
            g[f"states_L{L}_n{n}"] = np.asarray(b.get_states(n), dtype=np.int64)
    g["states_L10_n5"] = np.asarray(Basis(10).get_states(5), dtype=np.int64)
    g["states_types"] = np.array([type(Basis(4).get_states(n)).__name__ for n in (None, 0, 1, 2)])

    # 2. projector triplet streams (operators.py:305-527), exact order
    def trip(gen):
        r, c, v = [], [], []
        for i, j, val in gen:
            r.append(int(i)); c.append(int(j)); v.append(float(val))
        return np.asarray(r, np.int64), np.asarray(c, np.int64), np.asarray(v, np.float64)

    sec = Basis(4).get_sector(2, 2)
    up, dn = sec.up_states, sec.dn_states
    for name, gen in [
        ("hop_L4_22_03", ops.project_hopping(up, dn, 4, 0, 3, 1.0)),
        ("hop_L4_22_12_t07", ops.project_hopping(up, dn, 4, 1, 2, 0.7)),
        ("hop_L4_22_03_w0", ops.project_hopping(up, dn, 0, 0, 3, 1.0)),
        ("hop_L4_22_03_w2", ops.project_hopping(up, dn, 2, 0, 3, 1.0)),
        ("inter_L4_22_u4", ops.project_hubbard_inter(up, dn, np.full(4, 4.0))),
        ("inter_L4_22_uvar", ops.project_hubbard_inter(up, dn, np.array([1.0, 0.0, 2.5, 0.3]))),
        ("onsite_L4_22", ops.project_onsite_energy(up, dn, np.array([0.1, 0.2, 0.3, 0.4]))),
        ("onsite_L4_22_zero", ops.project_onsite_energy(up, dn, np.array([0.0, 0.0, 0.3, 0.0]))),
    ]:
        r, c, v = trip(gen)
        g[name + "_r"], g[name + "_c"], g[name + "_v"] = r, c, v
    sec = Basis(5).get_sector(3, 1)
    r, c, v = trip(ops.project_hopping(sec.up_states, sec.dn_states, 5, 1, 4, -0.5))
    g["hop_L5_31_14_r"], g["hop_L5_31_14_c"], g["hop_L5_31_14_v"] = r, c, v

    # 3. full model triplet streams (models/hubbard.py:13-22, anderson.py:147-158)
    models = {
        "hub_chain4_22": (HubbardModel(4, chain(4), inter=4.0, mu=2.0, hop=1.0), 2, 2),
        "hub_ring4_22": (HubbardModel(4, chain(4, True), inter=4.0, mu=2.0, hop=1.0), 2, 2),
        "hub_2x2_22": (HubbardModel(4, [[0, 1], [0, 2], [1, 3], [2, 3]], inter=4.0, mu=2.0, hop=1.0), 2, 2),
        "hub_chain5_32": (HubbardModel(5, chain(5), inter=3.0, eps=0.25, mu=1.0, hop=-0.8), 3, 2),
        "hub_chain6_33": (HubbardModel(6, chain(6), inter=4.0, mu=2.0, hop=1.0), 3, 3),
        "hub_ring6_33": (HubbardModel(6, chain(6, True), inter=4.0, mu=2.0, hop=1.0), 3, 3),
        "hub_chain3_10": (HubbardModel(3, chain(3), inter=4.0, mu=2.0, hop=1.0), 1, 0),
        "siam4_22": (SingleImpurityAndersonModel(u=2.0, eps_imp=0.0, eps_bath=[0.1, 0.2, 0.3],
                                                 v=[1.0, 0.7, 0.4]), 2, 2),
        "siam2_11": (SingleImpurityAndersonModel(u=4.0, v=[1.0], mu=2.0, eps_bath=0.0), 1, 1),
    }
    for name, (model, nu, nd) in models.items():
        s = model.get_sector(nu, nd)
        r, c, v = trip(model._hamiltonian_data(s.up_states, s.dn_states))
        g[name + "_r"], g[name + "_c"], g[name + "_v"] = r, c, v
        ham = model.hamiltonian(nu, nd)
        g[name + "_e0"] = np.array(np.linalg.eigvalsh(ham)[0])
        x = np.cos(0.37 * np.arange(ham.shape[0]))
        g[name + "_hv"] = model.hamilton_operator(nu, nd).matvec(x)
    g["hub2_11_ham"] = HubbardModel(2, [[0, 1]], inter=2.0, eps=1.0, hop=1.0).hamiltonian(1, 1)

    # 4. H.v + E0 at L=8 (4,4): HamiltonOperator._matvec (operators.py:626-630)
    m8 = HubbardModel(8, chain(8), inter=4.0, mu=2.0, hop=1.0)
    hop8 = m8.hamilton_operator(4, 4)
    x8 = np.cos(0.37 * np.arange(4900))
    g["hub_chain8_44_hv"] = hop8.matvec(x8)
    ham8 = hop8.toarray()
    ev8, evec8 = np.linalg.eigh(ham8)
    g["hub_chain8_44_e0"] = np.array(ev8[0])
    g["hub_chain8_44_trace"] = np.array(hop8.trace())
    # every sector of L=4 (incl. empty / full strings): H.v on cos vector
    m4 = HubbardModel(4, chain(4), inter=4.0, mu=2.0, hop=1.0)
    for nu in range(5):
        for nd in range(5):
            h = m4.hamilton_operator(nu, nd)
            x = np.cos(0.37 * np.arange(h.shape[0]))
            g[f"hub_chain4_all_{nu}{nd}_hv"] = h.matvec(x)

    # 5. Heisenberg (models/heisenberg.py:19-40)
    for N in (4, 6, 8, 10):
        hm = HeisenbergModel(refshim.ChainStandIn(N), j=1.0, jz=1.0)
        ham = hm.hamiltonian(s=0)
        g[f"heis_chain{N}_s0_e0"] = np.array(np.linalg.eigvalsh(ham)[0])
        if N <= 6:
            r, c, v = trip(hm._hamiltonian_data(hm.get_states(0)))
            g[f"heis_chain{N}_s0_r"], g[f"heis_chain{N}_s0_c"], g[f"heis_chain{N}_s0_v"] = r, c, v
        x = np.cos(0.37 * np.arange(ham.shape[0]))
        g[f"heis_chain{N}_s0_hv"] = hm.hamilton_operator(s=0).matvec(x)
    hm = HeisenbergModel(refshim.ChainStandIn(6), j=1.0, jz=0.0)
    g["heis_xx6_s0_e0"] = np.array(np.linalg.eigvalsh(hm.hamiltonian(s=0))[0])
    hm = HeisenbergModel(refshim.ChainStandIn(6, periodic=True), j=0.8, jz=1.3)
    g["heis_ring6_s1_ham"] = hm.hamiltonian(s=1)
    for N in (3, 4):
        hm = HeisenbergModel(refshim.ChainStandIn(N), j=1.0, jz=1.0)
        g[f"heis_chain{N}_full_ham"] = hm.hamiltonian()

    # 6. ladder operators, sigma=UP (operators.py:652-791); DN is broken in the reference
    b4 = Basis(4)
    for nu in range(4):
        for nd in range(5):
            s = b4.get_sector(nu, nd)
            s1 = b4.upper_sector(nu, nd, UP)
            for pos in range(4):
                cd = ops.CreationOperator(s, s1, pos, UP)
                x = np.cos(0.37 * np.arange(s.size)) + 0.5
                y = cd.matvec(x)
                g[f"cdag_L4_{nu}{nd}_p{pos}"] = y
                c = ops.AnnihilationOperator(s1, s, pos, UP)
                x1 = np.cos(0.21 * np.arange(s1.size)) + 0.5
                g[f"c_L4_{nu + 1}{nd}_p{pos}"] = c.matvec(x1)

    # 7. zero-temperature Lehmann G(z) from reference parts
    z = np.linspace(-6, 6, 1001) + 0.05j
    g["z_grid"] = z
    for L, pos in ((4, 0), (6, 0), (6, 2), (8, 0)):
        n = L // 2
        model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
        basis = model.basis
        s = basis.get_sector(n, n)
        ev, evec = ed.solve_sector(model, s)
        e0, gs = ev[0], evec[:, 0]
        sp = basis.upper_sector(n, n, UP)
        sm = basis.lower_sector(n, n, UP)
        evp, evecp = ed.solve_sector(model, sp)
        evm, evecm = ed.solve_sector(model, sm)
        cd_gs = ops.CreationOperator(s, sp, pos, UP).matvec(gs)
        c_gs = ops.AnnihilationOperator(s, sm, pos, UP).matvec(gs)
        wp = np.abs(evecp.T @ cd_gs) ** 2
        wm = np.abs(evecm.T @ c_gs) ** 2
        G = (wp[None, :] / (z[:, None] - evp[None, :] + e0)).sum(1)
        G += (wm[None, :] / (z[:, None] + evm[None, :] - e0)).sum(1)
        g[f"gf0T_chain{L}_p{pos}"] = G
        g[f"gf0T_chain{L}_p{pos}_e0"] = np.array(e0)
        g[f"gf0T_chain{L}_p{pos}_norms"] = np.array([cd_gs @ cd_gs, c_gs @ c_gs])
        print("zeroT", L, pos, time.time() - t00, flush=True)

    # 8. finite-T gf_lehmann (exactdiag.py:215-245)
    for L, beta in ((2, 10.0), (3, 10.0), (4, 10.0), (4, 50.0)):
        model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
        d = ed.gf_lehmann(model, z, beta=beta, pos=0, sigma=UP, occ=True)
        g[f"gfT_chain{L}_b{int(beta)}"] = d.gf
        g[f"gfT_chain{L}_b{int(beta)}_meta"] = np.array([d.gs_energy, d.occ, d.occ_double])
    siam = SingleImpurityAndersonModel(u=4.0, v=[1.0], mu=2.0, eps_bath=0.0, temp=0.1)
    g["gfT_siam2_b10"] = siam.impurity_gf(z)

    # 9. gf0_lehmann (greens.py:18-64)
    for L in (2, 3, 4, 5):
        ham0 = np.zeros((L, L))
        for i in range(L - 1):
            ham0[i, i + 1] = ham0[i + 1, i] = 1.0
        g[f"gf0_chain{L}"] = gf0_lehmann(ham0, z=z)

    # 10. reference Lanczos (exactdiag.py:324-375), global RNG seeded
    ham6 = HubbardModel(6, chain(6), inter=4.0, mu=2.0, hop=1.0).hamiltonian(3, 3)
    np.random.seed(1234)
    a, b = ed.lanczos_coeffs(ham6, 12)
    g["lanczos_ref_a"], g["lanczos_ref_b"] = np.asarray(a), np.asarray(b)
    np.random.seed(1234)
    g["lanczos_ref_psi0"] = np.random.uniform(0, 1, size=len(ham6))
    e_gs, vec = ed.lanczos_ground_state(a, b)
    g["lanczos_ref_egs"] = np.array(e_gs)

    # 11. compute_groundstate (exactdiag.py:24-43)
    for L in (2, 3, 4):
        model = HubbardModel(L, chain(L), inter=4.0, mu=2.0, hop=1.0)
        gs = ed.compute_groundstate(model)
        g[f"groundstate_chain{L}"] = np.array([gs.energy, gs.n_up, gs.n_dn])

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays", os.path.getsize(OUT) / 1e6, "MB", time.time() - t00, "s")


if __name__ == "__main__":
    main()
