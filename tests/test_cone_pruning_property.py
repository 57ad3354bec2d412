"""CPU property test behind the cone pipeline (scone_model.cu: cone_*_mb): with the oracle's dense fp64
formulation, zeroing every row of H_l outside  LIVE_l = (receptive cone of the readout) & (structural support of the flows)
after every layer changes neither the log-probs nor the weight gradients.

  cone:     C_L = edges incident to the neighbours of the last node (the rows `Bconds_func` reads,
            trajectory_experiments.py:298-303);  C_{l-1} = one hop around C_l in the pattern of I + L_lower + L_upper
  support:  S_0 = edges of the flow;  S_l = one hop around S_{l-1}   (no bias, act(0) = 0)
"""
import numpy as np
import pytest
import torch

from golden_util import Dataset
from oracle import scone_oracle as so


def _live_masks(S0, S1, B1, nbrhoods, last_nodes, flows, n_layers):
    E = S0.shape[0]
    A = ((np.abs(S0) + np.abs(S1) + np.eye(E)) > 0).astype(np.float64)         # merged operator pattern, symmetric
    B = len(last_nodes)
    cone = np.zeros((n_layers, B, E), bool)
    for t in range(B):
        nb = nbrhoods[last_nodes[t]]
        nb = nb[nb >= 0]
        cone[n_layers - 1, t] = (np.abs(B1[nb]).sum(axis=0) > 0)
    for l in range(n_layers - 2, -1, -1):
        cone[l] = (cone[l + 1].astype(np.float64) @ A) > 0
    supp = np.zeros((n_layers, B, E), bool)
    cur = flows.reshape(B, E) != 0
    for l in range(n_layers):
        cur = (cur.astype(np.float64) @ A) > 0
        supp[l] = cur
    return cone & supp, cone, supp


@pytest.mark.parametrize('model', ['scone', 'ebli'])
def test_live_row_pruning_is_exact(model):
    ds = Dataset('dataset_small.npz')
    S0, S1 = so.shift_matrices(ds.B1, ds.B2, model)
    nbrhoods, _, B1_jax = so.neighbourhood_tables(np.asarray(ds.B1), ds.last_nodes)
    L, C = 3, 8
    live, cone, supp = _live_masks(np.asarray(S0), np.asarray(S1), np.asarray(ds.B1), nbrhoods, ds.last_nodes, np.asarray(ds.flows), L)
    assert live.sum() < cone.sum() and live.sum() < supp.sum()                   # both prunings remove rows on this dataset
    dt = torch.float64
    rs = np.random.RandomState(5)
    shapes = [(1, C)] * 3 + [(C, C)] * 6 + [(C, 1)]
    S0t, S1t = torch.as_tensor(np.asarray(S0), dtype=dt), torch.as_tensor(np.asarray(S1), dtype=dt)
    X = torch.as_tensor(np.asarray(ds.flows), dtype=dt).reshape(len(ds.last_nodes), -1, 1)
    Bc = torch.as_tensor(B1_jax, dtype=dt)[torch.as_tensor(nbrhoods)[torch.as_tensor(np.asarray(ds.last_nodes), dtype=torch.int64)]]
    y = torch.as_tensor(np.asarray(ds.targets), dtype=dt).reshape(len(ds.last_nodes), -1, 1)
    act = torch.tanh if model == 'scone' else (lambda z: torch.where(z >= 0, z, 0.01 * z))

    def run(masks):
        W = [torch.as_tensor(0.3 * rs_w, dtype=dt).requires_grad_(True) for rs_w in Wnp]
        cur = X
        for i in range(L):
            cur = act(cur @ W[3 * i] + (S0t @ cur) @ W[3 * i + 1] + (S1t @ cur) @ W[3 * i + 2])
            if masks is not None:
                cur = cur * torch.as_tensor(masks[i], dtype=dt).unsqueeze(-1)
        logits = (Bc @ cur) @ W[-1]
        lp = logits - torch.logsumexp(logits, dim=1, keepdim=True)
        loss = -(lp * y).sum() / len(y)
        return lp.detach().numpy(), [g.numpy() for g in torch.autograd.grad(loss, W)]

    Wnp = [rs.randn(*s) for s in shapes]
    lp_full, g_full = run(None)
    lp_live, g_live = run(live)
    assert np.abs(lp_full - lp_live).max() <= 1e-13
    for a, b in zip(g_full, g_live):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())
    # the cone alone and the support alone are exact too (pipeline 2 = support; the first cone version = cone)
    for masks in (cone, supp):
        lp_m, g_m = run(masks)
        assert np.abs(lp_full - lp_m).max() <= 1e-13
        for a, b in zip(g_full, g_m):
            assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())
