"""Helpers to materialise the committed golden fixtures (tests/golden/*.npz, produced by
oracle/make_golden.py from the reference's own code) as the dense arrays the reference uses."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def dense_from_nz(nz, val, shape, dtype=np.float64):
    M = np.zeros(shape, dtype=dtype)
    M[nz[0], nz[1]] = val
    return M


class Dataset:
    """Dense view of dataset_*.npz in the reference's folder format (synthetic_data_gen.py:11-31)."""

    def __init__(self, name):
        d = load(name)
        self.raw = d
        self.N = int(d['n_nodes'])
        self.edges = d['edges']
        self.faces = d['faces']
        self.E, self.F = len(self.edges), len(self.faces)
        self.D = int(d['max_degree'])
        self.B1 = dense_from_nz(d['B1_nz'], d['B1_val'], (self.N, self.E))
        self.B2 = dense_from_nz(d['B2_nz'], d['B2_val'], (self.E, self.F))
        self.n_traj = len(d['last_nodes'])
        self.flows = dense_from_nz(d['flows_nz'], d['flows_val'], (self.n_traj, self.E)).reshape(self.n_traj, self.E, 1)
        if 'rev_flows_nz' in d.files:
            self.rev_flows = dense_from_nz(d['rev_flows_nz'], d['rev_flows_val'], (self.n_traj, self.E)).reshape(self.n_traj, self.E, 1)
        self.targets = np.zeros((self.n_traj, self.D, 1))
        self.targets[np.arange(self.n_traj), d['targets_argmax'], 0] = 1.0
        self.train_mask = d['train_mask'].astype(np.int64)
        self.test_mask = d['test_mask'].astype(np.int64)
        self.last_nodes = d['last_nodes'].astype(np.int64)
        self.target_nodes = d['target_nodes'].astype(np.int64)

    def shift(self, name):
        d = self.raw
        return dense_from_nz(d[name + '_nz'], d[name + '_val'], (self.E, self.E))


def weights_of(fx, tag):
    return [fx['%s_%d' % (tag, i)] for i in range(int(fx['n_weights']))]
