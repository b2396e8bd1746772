"""CPU suite of TA2N's soft-DTW (models/OTAM.py): the numpy restatement against the goldens
written from the reference's own CPU path (numba compute_softdtw / compute_softdtw_backward, SoftDTW module)."""
import numpy as np
import pytest
import torch

from oracle import clipspm_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", list(H.SOFTDTW_CASES))
def test_oracle_softdtw_matches_reference_golden(name):
    B, N, M, d, gamma, bw, seed = H.SOFTDTW_CASES[name]
    g = H.golden(name)
    X, Y, D = O.make_softdtw_inputs(B, N, M, d, seed)
    R = O.softdtw_forward_np(D.numpy(), gamma, bw)
    E = O.softdtw_backward_np(D.numpy(), R, gamma, bw)
    Rg = g["R"].numpy()
    fin = np.isfinite(Rg)
    assert (np.isfinite(R) == fin).all()
    assert np.abs(R[fin] - Rg[fin]).max() < 1e-5 * np.abs(Rg[fin]).max()
    assert np.abs(E - g["E"].numpy()).max() < 1e-5 * max(1.0, np.abs(g["E"].numpy()).max())
    assert H.rel_err(O.softdtw_module(X, Y, gamma, False, bw), g["module"]) < 1e-5
    k = min(N, M)
    assert H.rel_err(O.softdtw_module(X[:, :k], Y[:, :k], gamma, True, bw), g["module_norm"]) < 1e-4


def test_softdtw_hand_case():
    """1x1 and 2x2 tables by hand: R[1,1] = D[0,0]; gamma -> 0 recovers hard DTW (min over the three predecessors)."""
    D = np.array([[[0.3]]])
    assert abs(O.softdtw_forward_np(D, 1.0)[0, 1, 1] - 0.3) < 1e-12
    D = np.array([[[1.0, 5.0], [2.0, 1.5]]])
    R = O.softdtw_forward_np(D, 1e-3)
    assert abs(R[0, 2, 2] - (1.0 + 1.5)) < 1e-2       # diagonal path
    E = O.softdtw_backward_np(D, R, 1e-3)
    assert abs(E[0, 0, 0] - 1) < 1e-6 and abs(E[0, 1, 1] - 1) < 1e-6 and E[0, 0, 1] < 1e-6
