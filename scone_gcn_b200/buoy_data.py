"""Ocean-drifter dataset converter — host-side mirror of ocean_drifters_data/buoy_data.py (Schaub's Madagascar hex grid).

    convert_drifters('dataBuoys.jld2', folder_suffix='buoy')   # writes trajectory_data_{1,2}hop_buoy/ (+ prefixes.npy)

Same steps as the reference script (buoy_data.py:20-136): 0-index the edge / triangle lists, build B1 / B2, strip
back-and-forth steps, keep the last 10 nodes of every path with >= 5 nodes, seed-1 80/20 split, untruncated 1-hop / 2-hop
datasets for the forward and reversed paths.  The JLD2 file is read with scone_gcn_b200.jld2 (h5py is not needed).
"""
import os
import pickle

import numpy as np

from . import synthetic_data_gen as sdg
from .jld2 import JLD2File


def read_drifter_file(path):
    """(edge_list [2,E] 0-indexed, face_list [3,F] 0-indexed, traj_nodes list of 0-indexed node lists) — buoy_data.py:20-36."""
    f = JLD2File(path)
    edge_list = np.asarray(f.read('elist')) - 1
    face_list = np.asarray(f.read('tlist')) - 1
    traj_nodes = [[int(x) - 1 for x in traj] for traj in f.read('TrajectoriesNodes')]
    return edge_list, face_list, traj_nodes


def build_drifter_dataset(edge_list, face_list, traj_nodes):
    """Everything buoy_data.py:38-100 computes, as a dict of arrays (no files written)."""
    import networkx as nx
    G = nx.Graph()
    G.add_edges_from([(int(edge_list[0][i]), int(edge_list[1][i])) for i in range(len(edge_list[0]))])
    V, E = np.array(sorted(G.nodes)), np.array([sorted(x) for x in sorted(G.edges)])
    faces = np.array(sorted([[int(face_list[j][i]) for j in range(3)] for i in range(len(face_list[0]))]))
    edge_to_idx = {tuple(int(v) for v in e): i for i, e in enumerate(E)}
    B1, B2 = sdg.incidence_matrices(G, V, E, faces, edge_to_idx)
    G_undir = G.to_undirected()
    paths = [path[-10:] for path in sdg.strip_paths(traj_nodes) if len(path) >= 5]
    rev_paths = [path[::-1] for path in paths]
    np.random.seed(1)
    train_mask = np.asarray([1] * round(len(paths) * 0.8) + [0] * round(len(paths) * 0.2))
    np.random.shuffle(train_mask)
    test_mask = 1 - train_mask
    max_degree = np.max([deg for _, deg in G_undir.degree()])
    fwd = sdg.path_dataset(G_undir, E, edge_to_idx, paths, max_degree, include_2hop=True, truncate_paths=False)
    rev = sdg.path_dataset(G_undir, E, edge_to_idx, rev_paths, max_degree, include_2hop=True, truncate_paths=False)
    return dict(G_undir=G_undir, V=V, E=E, faces=faces, B1=B1, B2=B2, paths=paths, train_mask=train_mask, test_mask=test_mask,
                max_degree=int(max_degree), fwd=fwd, rev=rev, prefixes=[path[:-2] for path in paths])


FILENAMES = ('flows_in', 'B1', 'B2', 'targets', 'train_mask', 'test_mask', 'G_undir', 'last_nodes', 'target_nodes', 'rev_flows_in',
             'rev_targets', 'rev_last_nodes', 'rev_target_nodes')


def convert_drifters(jld2_path, folder_suffix='buoy', out_dir='.'):
    """Writes the reference folder format (buoy_data.py:102-136) and returns the dataset dict."""
    d = build_drifter_dataset(*read_drifter_file(jld2_path))
    fwd, rev = d['fwd'], d['rev']
    sets = {1: [fwd[0], d['B1'], d['B2'], fwd[1], d['train_mask'], d['test_mask'], d['G_undir'], fwd[2], fwd[3], rev[0], rev[1], rev[2], rev[3]],
            2: [fwd[4], d['B1'], d['B2'], fwd[5], d['train_mask'], d['test_mask'], d['G_undir'], fwd[6], fwd[7], rev[4], rev[5], rev[6], rev[7]]}
    for hop, arrs in sets.items():
        folder = os.path.join(out_dir, 'trajectory_data_%dhop_%s' % (hop, folder_suffix))
        os.makedirs(folder, exist_ok=True)
        for arr, name in zip(arrs, FILENAMES):
            if name == 'G_undir':
                with open(os.path.join(folder, name + '.pkl'), 'wb') as f:
                    pickle.dump(d['G_undir'], f, protocol=4)
            else:
                np.save(os.path.join(folder, name + '.npy'), arr)
    np.save(os.path.join(out_dir, 'trajectory_data_1hop_%s' % folder_suffix, 'prefixes.npy'),
            np.array(d['prefixes'], dtype=object), allow_pickle=True)
    return d
