"""Host-side data code (scone_gcn_b200.synthetic_data_gen, complex index lists) against the reference's own
output (tests/golden/dataset_*.npz were written by the reference generator).  CPU only."""
import os

import numpy as np
import pytest

from golden_util import Dataset
from scone_gcn_b200 import synthetic_data_gen as sdg
from scone_gcn_b200.complex import flows_to_csr, incidence_lists_from_dense, incidence_lists_from_simplices


@pytest.mark.parametrize('n,m,name', [(400, 1000, 'dataset_default.npz'), (120, 60, 'dataset_small.npz')])
def test_generator_reproduces_reference_bit_for_bit(tmp_path, monkeypatch, n, m, name):
    ref = Dataset(name)
    monkeypatch.chdir(tmp_path)
    sdg.generate_dataset(n, m, 'x')
    X, (B1, B2), y, train_mask, test_mask, G_undir, last_nodes, target_nodes = sdg.load_dataset('trajectory_data_1hop_x')
    assert np.array_equal(B1, ref.B1) and np.array_equal(B2, ref.B2)
    assert np.array_equal(X, ref.flows) and np.array_equal(y, ref.targets)
    assert np.array_equal(train_mask, ref.train_mask) and np.array_equal(test_mask, ref.test_mask)
    assert np.array_equal(last_nodes, ref.last_nodes) and np.array_equal(target_nodes, ref.target_nodes)
    assert np.array_equal(np.load('trajectory_data_1hop_x/rev_flows_in.npy'), ref.rev_flows)
    assert sorted(os.listdir('trajectory_data_2hop_x')) == sorted(os.listdir('trajectory_data_1hop_x'))
    assert max(d for _, d in G_undir.degree()) == ref.D


def test_incidence_lists_two_routes_agree():
    ds = Dataset('dataset_default.npz')
    a = incidence_lists_from_dense(ds.B1, ds.B2)
    b = incidence_lists_from_simplices(ds.edges, ds.faces)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.all(a[3] == np.array([1, -1, 1]))           # (+,-,+) in ascending edge order


def test_sparse_dataset_round_trip():
    ds = Dataset('dataset_small.npz')
    sp = sdg.SparseDataset.from_dense(ds.flows, ds.B1, ds.B2, ds.targets, ds.train_mask, ds.test_mask, ds.last_nodes,
                                      ds.target_nodes)
    X, B1, B2, y = sp.to_dense()
    assert np.array_equal(X, ds.flows) and np.array_equal(B1, ds.B1) and np.array_equal(B2, ds.B2)
    assert np.array_equal(y, ds.targets)
    assert np.array_equal(sp.faces, ds.faces) and np.array_equal(sp.edges, ds.edges)


def test_fast_generator_invariants():
    sp = sdg.generate_sparse_dataset(2000, 200, seed=3, n_waypoints=8)
    N, E = int(sp.n_nodes), len(sp.edges)
    assert sp.n_traj == 200 and sp.train_mask.sum() == 160
    lut = {(int(a), int(b)): i for i, (a, b) in enumerate(sp.edges)}
    nbrs = [[] for _ in range(N)]
    for a, b in sp.edges:
        nbrs[a].append(int(b))
        nbrs[b].append(int(a))
    assert max(len(v) for v in nbrs) == int(sp.max_degree)
    for t in range(sp.n_traj):
        fe = sp.flow_edge[sp.traj_ptr[t]:sp.traj_ptr[t + 1]]
        fv = sp.flow_val[sp.traj_ptr[t]:sp.traj_ptr[t + 1]]
        assert np.all(np.diff(fe) > 0) and np.all(np.abs(fv) == 1)         # simple path: each edge once
        flow = np.zeros(E)
        flow[fe] = fv
        path = sdg.flow_to_path(flow, sp.edges, sp.last_nodes[t])            # must be one chain ending at last node
        assert len(path) == len(fe) + 1 and path[-1] == sp.last_nodes[t]
        nb = sorted(nbrs[sp.last_nodes[t]])
        assert nb[sp.target_idx[t]] == sp.target_nodes[t]
        assert np.array_equal(sdg.path_to_flow(path, lut, E)[:, 0], flow)


def test_flows_to_csr():
    X = np.zeros((3, 5, 1))
    X[0, 1], X[0, 4], X[2, 0] = 1, -1, 1
    ptr, e, v = flows_to_csr(X)
    assert ptr.tolist() == [0, 2, 2, 3] and e.tolist() == [1, 4, 0] and v.tolist() == [1, -1, 1]
