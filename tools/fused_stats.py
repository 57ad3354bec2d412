"""Distribution of the fused pipeline's per-trajectory plan sizes on a bench workload (GPU): live rows per layer, cone size."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_gcn_b200 as sg
from scone_gcn_b200 import synthetic_data_gen as sdg

n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 370000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sp = sdg.generate_sparse_dataset(n_nodes, B, seed=1030, n_waypoints=24)
cx = sg.SimplicialComplex.from_simplices(int(sp.n_nodes), sp.edges, sp.faces, 'scone')
net = sg.SconeModel(cx, [32, 32, 32], micro_batch=B)
rs = np.random.RandomState(1030)
net.set_weights([0.01 * rs.randn(*s) for s in net.shapes])
nnz = int(sp.traj_ptr[B])
net.forward(sp.traj_ptr[:B + 1], sp.flow_edge[:nnz], sp.flow_val[:nnz], sp.last_nodes[:B])
print('fused_info', net.fused_info())
H = np.stack([net.fused_header(t) for t in range(B)])
def pct(a):
    return ' '.join('%s=%d' % (k, np.percentile(a, q)) for k, q in (('p50', 50), ('p90', 90), ('p99', 99), ('p99.9', 99.9), ('max', 100))) + ' mean=%.1f' % a.mean()
for name, col in (('n1', 1), ('n2', 2), ('n3', 3), ('hash', 11), ('pairs', 10), ('flows', 13)):
    print(name, pct(H[:, col]))
tot = H[:, 1] + H[:, 2] + H[:, 3]
print('tot', pct(tot), 'frac tot>256: %.4f' % (tot > 256).mean(), 'flags nonzero:', int((H[:, 0] != 0).sum()))
print('nnz per traj', pct(np.diff(sp.traj_ptr[:B + 1])))
i = np.argsort(-H[:, 11])[:8]
print('largest cones: hash entries', H[i, 11], 'tot', tot[i], 'deg(last)', [(cx.nbrhoods[n] >= 0).sum() for n in sp.last_nodes[i]])
