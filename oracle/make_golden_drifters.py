#!/usr/bin/env python
"""tests/golden/dataset_drifters.npz: the ocean-drifter dataset as the REFERENCE's functions build it.

TEST INFRASTRUCTURE.  ocean_drifters_data/buoy_data.py is a run-on-import script that needs h5py (absent), so its steps
(:20-100) are replayed here verbatim on top of the reference's own synthetic_data_gen functions (strip_paths,
incidence_matrices, path_dataset — imported from /root/reference, unmodified); only the three raw arrays it pulls out of
dataBuoys.jld2 come from scone_gcn_b200.jld2 instead of h5py.  Run in the build container: python oracle/make_golden_drifters.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(1, '/root/reference/trajectory_analysis')
sys.path.insert(2, REPO)
import numpy as np  # noqa: E402
import networkx as nx  # noqa: E402
import compat  # noqa: E402

compat.apply()
import synthetic_data_gen as ref  # noqa: E402  (the reference module)
from scone_gcn_b200.jld2 import JLD2File  # noqa: E402

f = JLD2File('/root/reference/ocean_drifters_data/dataBuoys.jld2')
edge_list = np.asarray(f.read('elist')) - 1                                       # buoy_data.py:20
face_list = np.asarray(f.read('tlist')) - 1                                       # :23
traj_nodes = [[int(x) - 1 for x in t] for t in f.read('TrajectoriesNodes')]       # :36
G = nx.Graph()
G.add_edges_from([(edge_list[0][i], edge_list[1][i]) for i in range(len(edge_list[0]))])
V, E = np.array(sorted(G.nodes)), np.array([sorted(x) for x in sorted(G.edges)])
faces = np.array(sorted([[face_list[j][i] for j in range(3)] for i in range(len(face_list[0]))]))
edge_to_idx = {tuple(e): i for i, e in enumerate(E)}
B1, B2 = ref.incidence_matrices(G, V, E, faces, edge_to_idx)
G_undir = G.to_undirected()
paths = [path[-10:] for path in ref.strip_paths(traj_nodes) if len(path) >= 5]      # :55-57
np.random.seed(1)
train_mask = np.asarray([1] * round(len(paths) * 0.8) + [0] * round(len(paths) * 0.2))
np.random.shuffle(train_mask)
max_degree = np.max([deg for n, deg in G_undir.degree()])
fl1, tg1, ln1, sf1, fl2, tg2, ln2, sf2 = ref.path_dataset(G_undir, E, edge_to_idx, paths, max_degree, include_2hop=True,
                                                          truncate_paths=False)
X = fl1[:, :, 0]
out = dict(n_nodes=np.int64(B1.shape[0]), edges=E.astype(np.int32), faces=faces.astype(np.int32), max_degree=np.int64(max_degree),
           B1_nz=np.stack(np.nonzero(B1)).astype(np.int32), B1_val=B1[np.nonzero(B1)].astype(np.int8),
           B2_nz=np.stack(np.nonzero(B2)).astype(np.int32), B2_val=B2[np.nonzero(B2)].astype(np.int8),
           flows_nz=np.stack(np.nonzero(X)).astype(np.int32), flows_val=X[np.nonzero(X)].astype(np.int8),
           targets_argmax=np.argmax(tg1[:, :, 0], axis=1).astype(np.int32), last_nodes=np.asarray(ln1, np.int32),
           target_nodes=np.asarray(sf1, np.int32), train_mask=train_mask.astype(np.int8), test_mask=(1 - train_mask).astype(np.int8),
           path_ptr=np.cumsum([0] + [len(p) for p in paths]).astype(np.int32), path_nodes=np.concatenate(paths).astype(np.int32),
           raw_edge_list=edge_list.astype(np.int32), raw_face_list=face_list.astype(np.int32),
           raw_traj_ptr=np.cumsum([0] + [len(t) for t in traj_nodes]).astype(np.int32),
           raw_traj_nodes=np.concatenate([np.asarray(t, np.int32) for t in traj_nodes]),
           n_raw_traj=np.int64(len(traj_nodes)))
np.savez_compressed(os.path.join(REPO, 'tests', 'golden', 'dataset_drifters.npz'), **out)
print('nodes %d edges %d faces %d D %d raw traj %d usable %d train %d' % (B1.shape[0], B1.shape[1], B2.shape[1], max_degree,
                                                                       len(traj_nodes), len(paths), train_mask.sum()))
