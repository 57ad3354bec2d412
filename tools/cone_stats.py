"""Sizes of the receptive cone, the structural support and the LIVE rows (cone & support) per trajectory and layer — the sets the
cone pipeline builds on the device.  CPU only (index-only handle).  Usage: python tools/cone_stats.py [n_nodes] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

import scone_gcn_b200 as sg
from scone_gcn_b200 import synthetic_data_gen as sdg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L = 3
d = sdg.generate_sparse_dataset(n, B, seed=1030, n_waypoints=8)
cx = sg.SimplicialComplex.from_simplices(int(d.n_nodes), d.edges, d.faces, 'scone', index_only=True)
E = cx.E
P = None                                                    # merged operator pattern: S0 | S1 | I
for k in range(2):
    rp, col, _ = cx.shift_csr(k)
    M = sp.csr_matrix((np.ones(len(col), np.int8), col, rp), shape=(E, E))
    P = M if P is None else P + M
P = ((P + sp.identity(E, dtype=np.int8, format='csr')) > 0).astype(np.int8).tocsr()
nb = cx.nbrhoods
inc = sp.csr_matrix((np.ones(2 * E, np.int8), (np.r_[d.edges[:, 0], d.edges[:, 1]], np.r_[np.arange(E), np.arange(E)])),
                    shape=(cx.N, E))                      # node -> incident edges
top = np.zeros((E, B), np.int8)
for t in range(B):
    nbrs = nb[d.last_nodes[t]]
    nbrs = nbrs[nbrs >= 0]
    top[np.unique(inc[nbrs].indices), t] = 1
cone = [None] * L
cone[L - 1] = top > 0
for l in range(L - 2, -1, -1):
    cone[l] = (P @ cone[l + 1].astype(np.int8)) > 0
X = np.zeros((E, B), np.int8)
X[d.flow_edge[:d.traj_ptr[B]], np.repeat(np.arange(B), np.diff(d.traj_ptr[:B + 1]))] = 1
supp, cur = [], X
for l in range(L):
    cur = ((P @ cur) > 0).astype(np.int8)
    supp.append(cur > 0)
print('E = %d, %d trajectories, mean flow entries %.1f' % (E, B, X.sum() / B))
for l in range(L):
    live = cone[l] & supp[l]
    print('H_%d: cone %.1f  support %.1f  live %.1f rows per trajectory' % (l + 1, cone[l].sum() / B, supp[l].sum() / B, live.sum() / B))
