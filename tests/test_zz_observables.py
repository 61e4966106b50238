"""GPU parity of the spin observables (cmpy_b200/observables.py)
against the restatement of scripts/heisenberg.py's helpers in oracle/oracle_np.py."""
import numpy as np
import pytest

import oracle_np as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import cmpy_b200

    return cmpy_b200


@pytest.mark.parametrize("N,s", [(8, 0), (10, 1), (12, 0), (18, 0)])
def test_sz_observables_vs_oracle(cm, N, s):
    from cmpy_b200 import observables as obs

    states = orc.spin_states(N, s)
    rng = np.random.default_rng(N)
    gs = rng.standard_normal(len(states))
    gs /= np.linalg.norm(gs)
    for pos in (0, N // 2, N - 1):
        assert abs(obs.sz_expval(N, s, gs, pos) - orc.sz_expval(states, gs, pos)) < 1e-13
    for delta in (1, 2, N - 1):
        assert abs(obs.sz_correl(N, s, gs, delta, j=1.3) - orc.sz_correl(states, gs, delta, j=1.3)) < 1e-13
    assert abs(obs.sz_correl(N, s, gs, 3, pos=2) - orc.sz_correl(states, gs, 3, pos=2)) < 1e-13
    corr = obs.spin_correlations(N, s, gs, pos=1)
    ref = [0.25 if k == 1 else orc.sz_correl(states, gs, abs(k - 1), pos=min(k, 1)) for k in range(N)]
    assert np.abs(corr - np.asarray(ref)).max() < 1e-13
    # sum rule: sum_k <Sz_1 Sz_k> = s <Sz_1>
    assert abs(corr.sum() - s * obs.sz_expval(N, s, gs, 1)) < 1e-12
    with pytest.raises(ValueError):
        obs.sz_expval(N + 1, 0, gs, 0)


def test_ground_state_correlations_heisenberg_chain(cm):
    """Antiferromagnetic chain: the ground-state correlations alternate in sign, and the bond
    correlators add up to the expectation value of the model's own diagonal (in the reference
    scaling a bond listed in both directions contributes 2 jz Sz_i Sz_j, cmpy/models/heisenberg.py:28-31)."""
    from cmpy_b200.exactdiag import lanczos_run
    from cmpy_b200.models import HeisenbergModel
    from cmpy_b200 import observables as obs
    from refshim import ChainStandIn

    N, jz = 12, 1.0
    model = HeisenbergModel(ChainStandIn(N), j=1.0, jz=jz)
    h = model.hamilton_operator(s=0)
    res = lanczos_run(h, None, maxit=400, tol=1e-13, resid_tol=1e-10, want_vector=True)
    assert abs(res.e0 - (-5.903591587651)) < 1e-9
    gs = np.asarray(res.vector.cpu().numpy() if hasattr(res.vector, "cpu") else res.vector, dtype=np.float64)
    gs = gs / np.linalg.norm(gs)
    c = obs.spin_correlations(N, 0, gs, pos=0)
    assert abs(c[0] - 0.25) < 1e-14 and c[1] < 0 < c[2] and c[3] < 0
    zz = sum(obs.sz_correl(N, 0, gs, 1, pos=i) for i in range(N - 1))
    assert abs(2.0 * jz * zz - float(np.dot(gs * gs, h.diagonal()))) < 1e-12
