"""How sparse is the real workload at tile granularity? (CPU, index-only handle.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import scone_gcn_b200 as sg
from scone_gcn_b200 import synthetic_data_gen as sdg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 110000
B = 64
d = sdg.generate_sparse_dataset(n, B, seed=1030, n_waypoints=8)
E = len(d.edges)
for reorder in ('0', '1'):
    os.environ['SCONE_B200_NO_REORDER'] = '1' if reorder == '0' else '0'
    cx = sg.SimplicialComplex.from_simplices(int(d.n_nodes), d.edges, d.faces, 'scone', index_only=True)
    rank = cx.edge_rank
    P = None
    for k in range(2):
        rp, col, val = cx.shift_csr(k)
        M = sp.csr_matrix((np.ones(len(col), np.int8), col, rp), shape=(E, E))
        P = M if P is None else (P + M)
    P = (P + sp.identity(E, dtype=np.int8, format='csr')).tocsr()
    X = np.zeros((E, B), np.int8)
    X[d.flow_edge, np.repeat(np.arange(B), np.diff(d.traj_ptr))] = 1
    occ = X
    for layer in range(3):
        occ = (P @ occ) > 0                     # support of H_{layer+1}
        o = np.zeros_like(occ); o[rank] = occ   # internal row order
        TE, TT = 32, 4
        Ep = (E + TE - 1) // TE * TE
        pad = np.zeros((Ep, B), bool); pad[:E] = o
        tiles = pad.reshape(Ep // TE, TE, B // TT, TT).any(axis=(1, 3))
        units = pad.reshape(Ep, B // TT, TT).any(axis=2)
        print('reorder=%s H%d: rows active %.3f%%  units(edge x 4traj) %.3f%%  tiles(32e x 4t) non-empty %.2f%%  active rows per non-empty tile %.1f'
              % (reorder, layer + 1, 100 * occ.mean(), 100 * units.mean(), 100 * tiles.mean(), occ.sum() / max(tiles.sum(), 1)))
