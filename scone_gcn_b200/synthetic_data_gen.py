"""Dataset format + generators — host-side mirror of trajectory_analysis/synthetic_data_gen.py.

Two layers:
  1. The reference's folder format and small-scale generator (synthetic_data_gen.py:11-31,82-447), restated so that
     it runs on current NumPy / NetworkX and reproduces the reference's arrays bit-for-bit under the same seeds
     (tests/test_data_format.py checks this against tests/golden/dataset_default.npz, which the reference itself
     produced).  Same function names and argument meaning.
  2. A sparse dataset (`SparseDataset`) for complexes whose dense B1 / flows cannot exist (configs 4-5: a
     335k x 1M float64 B1 is 2.7 TB) plus a fast generator (`generate_sparse_dataset`) following the same recipe
     (uniform points sorted by x+y, Delaunay, two disc holes, BEGIN -> A_k -> B_k -> END shortest-path walks)
     with batched BFS trees instead of 3 x nx.shortest_path per walk.  `SparseDataset.to_dense()` gives the
     reference format back (round-trip tested at small scale).
Plotting (color_faces) and the RNN exporter (to_rnn_format) are outside the accelerated path and not provided.
"""
import os
import pickle

import numpy as np

# ---------------------------------------------------------------------------------------------------------------
# 1. reference format, small scale
# ---------------------------------------------------------------------------------------------------------------


def strip_paths(paths):
    """Remove back-and-forth steps a->b->a (synthetic_data_gen.py:43-61)."""
    out = []
    for path in paths:
        keep = []
        for node in path:
            if len(keep) >= 2 and node == keep[-2]:
                keep.pop()
            else:
                keep.append(node)
        out.append(keep)
    return out


def _complex_arrays(n, holes=True):
    """coords (sorted by x+y), valid node ids, faces [F,3], edges [E,2] — synthetic_data_gen.py:98-127."""
    from scipy.spatial import Delaunay
    np.random.seed(1)
    coords = np.random.rand(n, 2)
    coords = coords[np.argsort(np.sum(coords, axis=1))]
    np.random.seed(1030)
    tri = Delaunay(coords)
    keep = (np.linalg.norm(coords - [1 / 4, 3 / 4], axis=1) > 1 / 8) & (np.linalg.norm(coords - [3 / 4, 1 / 4], axis=1) > 1 / 8)
    valid_idxs = np.where(keep)[0] if holes else np.arange(n)
    simp = np.sort(tri.simplices, axis=1)
    simp = simp[keep[simp].all(axis=1)] if holes else simp
    faces = np.unique(simp, axis=0)                                        # lexicographic, as sorted(...) does
    sides = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [0, 2]]])
    edges = np.unique(sides, axis=0)
    return coords, valid_idxs, faces, edges


def random_SC_graph(n, holes=True):
    """Same return tuple as the reference (synthetic_data_gen.py:82-137): G, V, E, faces, edge_to_idx, coords, valid_idxs."""
    import networkx as nx
    coords, valid_idxs, faces, E = _complex_arrays(n, holes)
    G = nx.DiGraph()
    G.add_nodes_from(np.arange(n))
    for e in E:
        G.add_edge(*e)
    V = np.array(G.nodes)
    edge_to_idx = {tuple(E[i]): i for i in range(len(E))}
    print('Average degree:', np.average([G.degree[node] for node in range(n)]))
    print('Nodes:', len(V), 'Edges:', len(E))
    return G, V, E, faces, edge_to_idx, coords, valid_idxs


def incidence_matrices(G, V, E, faces, edge_to_idx):
    """Dense B1 (|V| x |E|), B2 (|E| x |faces|) with the reference's signs (synthetic_data_gen.py:139-161)."""
    E = np.asarray(E)
    nE, nV = len(E), len(V)
    pos = {int(v): i for i, v in enumerate(V)}
    B1 = np.zeros([nV, nE])
    B1[[pos[int(a)] for a in E[:, 0]], np.arange(nE)] = -1
    B1[[pos[int(b)] for b in E[:, 1]], np.arange(nE)] = 1
    B2 = np.zeros([nE, len(faces)])
    for f_idx, (a, b, c) in enumerate(np.asarray(faces)):
        B2[edge_to_idx[(a, b)], f_idx] = 1
        B2[edge_to_idx[(b, c)], f_idx] = 1
        B2[edge_to_idx[(a, c)], f_idx] = -1
    return B1, B2


def faces_from_B2(B2, E):
    out = []
    for j in range(B2.shape[1]):
        nodes = set()
        for e in np.asarray(E)[np.where(B2[:, j] != 0)]:
            nodes.update(int(v) for v in e)
        out.append(tuple(sorted(nodes)))
    return out


def _regions(points, valid_idxs):
    """BEGIN / END / A0-2 / B0-2 node sets (synthetic_data_gen.py:199-217)."""
    pv = points[valid_idxs]
    s = np.sum(pv, axis=1)
    BEGIN, END = valid_idxs[s < 1 / 4], valid_idxs[s > 7 / 4]

    def split(sel):
        d = points[sel, 1] - points[sel, 0]
        return sel[(d < 1 / 2) & (d > -1 / 2)], sel[d > 1 / 2], sel[d < -1 / 2]
    A = split(valid_idxs[(s > 1 / 4) & (s < 1)])
    B = split(valid_idxs[(s < 7 / 4) & (s > 1)])
    return BEGIN, END, A, B


def generate_random_walks(G, points, valid_idxs, m=1000):
    """m simple BEGIN -> A_k -> B_k -> END walks, k cycling 0,1,2 (synthetic_data_gen.py:178-243); same RNG draws."""
    import networkx as nx
    BEGIN, END, A, B = _regions(points, valid_idxs)
    paths = []
    G_undir = G.to_undirected()
    i = 0
    while len(paths) < m:
        v_begin = np.random.choice(BEGIN)
        v_1 = np.random.choice(A[i % 3])
        v_2 = np.random.choice(B[i % 3])
        v_end = np.random.choice(END)
        path = nx.shortest_path(G_undir, v_begin, v_1)[:-1] + nx.shortest_path(G_undir, v_1, v_2)[:-1] + \
            nx.shortest_path(G_undir, v_2, v_end)
        if len(path) == len(set(path)):
            paths.append(path)
            i += 1
    return G_undir, paths


def split_paths(paths, truncate_paths=True, suffix_size=2):
    """Truncate (random length) and split into prefix + suffix (synthetic_data_gen.py:245-258)."""
    if truncate_paths:
        paths = [p[:4 + np.random.choice(range(2, len(p) - 4))] for p in paths]
    prefixes = [p[:-suffix_size] for p in paths]
    suffixes = [p[-suffix_size:] for p in paths]
    return prefixes, suffixes, [p[-1] for p in prefixes]


def conditional_incidence_matrix(B1, Nv, D):
    B_cond = np.zeros([D, B1.shape[1]])
    B_cond[:len(Nv), :] = B1[Nv]
    return B_cond


def neighborhood(G, v):
    return np.array(sorted(G[v]))


def neighborhood_to_onehot(Nv, w, D):
    """One-hot over the sorted neighbours, zero-padded to D, as a column (synthetic_data_gen.py:288-297)."""
    out = np.zeros(D)
    out[:len(Nv)] = (np.asarray(Nv) == w).astype(float)
    return out.reshape(D, 1)


def flow_to_path(flow, E, last_node):
    """Inverse of path_to_flow for simple paths (synthetic_data_gen.py:299-325)."""
    flow = np.asarray(flow).reshape(-1)
    pred = {}
    for i in np.where(flow != 0)[0]:
        a, b = E[i]
        if flow[i] == 1:
            pred[int(b)] = int(a)
        elif flow[i] == -1:
            pred[int(a)] = int(b)
    path = [int(last_node)]
    while pred:
        if path[-1] not in pred:
            raise ValueError
        path.append(pred.pop(path[-1]))
    return path[::-1]


def path_to_flow(path, edge_to_idx, m):
    """+1 on edges walked low->high node id, -1 otherwise (synthetic_data_gen.py:327-344)."""
    f = np.zeros([m, 1])
    for v0, v1 in zip(path[:-1], path[1:]):
        if v0 < v1:
            f[edge_to_idx[(v0, v1)]] += 1
        else:
            f[edge_to_idx[(v1, v0)]] -= 1
    return f


def path_dataset(G_undir, E, edge_to_idx, paths, max_degree, include_2hop=True, truncate_paths=True):
    """1-hop and 2-hop flows / one-hot targets from paths (synthetic_data_gen.py:346-373)."""
    prefixes_1hop, suffixes, last_nodes = split_paths(paths, truncate_paths=truncate_paths,
                                                      suffix_size=(2 if include_2hop else 1))
    suffixes_1hop = [s[0] for s in suffixes]
    flows = np.array([path_to_flow(p, edge_to_idx, len(E)) for p in prefixes_1hop])
    targets = np.array([neighborhood_to_onehot(neighborhood(G_undir, p[-1]), s, max_degree)
                        for p, s in zip(prefixes_1hop, suffixes_1hop)])
    if not include_2hop:
        return flows, targets, last_nodes, suffixes_1hop, [], [], [], []
    prefixes_2hop = [np.concatenate([p, [s]]) for p, s in zip(prefixes_1hop, suffixes_1hop)]
    suffixes_2hop = [s[1] for s in suffixes]
    last_nodes_2hop = [s[0] for s in suffixes]
    flows_2hop = np.array([path_to_flow(p, edge_to_idx, len(E)) for p in prefixes_2hop])
    targets_2hop = np.array([neighborhood_to_onehot(neighborhood(G_undir, p[-1]), s, max_degree)
                             for p, s in zip(prefixes_2hop, suffixes_2hop)])
    return flows, targets, last_nodes, suffixes_1hop, flows_2hop, targets_2hop, last_nodes_2hop, suffixes_2hop


FILENAMES = ('flows_in', 'B1', 'B2', 'targets', 'train_mask', 'test_mask', 'G_undir', 'coords', 'last_nodes', 'target_nodes',
             'rev_flows_in', 'rev_targets', 'rev_last_nodes', 'rev_target_nodes')


def generate_dataset(n, m, folder, holes=True):
    """Writes trajectory_data_{1,2}hop_<folder>/ in the reference format (synthetic_data_gen.py:375-428)."""
    G, V, E, faces, edge_to_idx, coords, valid_idxs = random_SC_graph(n, holes=holes)
    B1, B2 = incidence_matrices(G, V, E, faces, edge_to_idx)
    G_undir, paths = generate_random_walks(G, coords, valid_idxs, m=m)
    rev_paths = [path[::-1] for path in paths]
    train_mask = np.asarray([1] * int(len(paths) * 0.8) + [0] * int(len(paths) * 0.2))
    np.random.shuffle(train_mask)
    test_mask = 1 - train_mask
    max_degree = np.max([deg for _, deg in G_undir.degree()])
    print(max_degree)
    fwd = path_dataset(G_undir, E, edge_to_idx, paths, max_degree)
    rev = path_dataset(G_undir, E, edge_to_idx, rev_paths, max_degree)
    sets = {1: [fwd[0], B1, B2, fwd[1], train_mask, test_mask, G_undir, coords, fwd[2], fwd[3], rev[0], rev[1], rev[2], rev[3]],
            2: [fwd[4], B1, B2, fwd[5], train_mask, test_mask, G_undir, coords, fwd[6], fwd[7], rev[4], rev[5], rev[6], rev[7]]}
    for hop, arrs in sets.items():
        d = 'trajectory_data_%dhop_%s' % (hop, folder)
        os.makedirs(d, exist_ok=True)
        for arr, name in zip(arrs, FILENAMES):
            if name == 'G_undir':
                with open(os.path.join(d, name + '.pkl'), 'wb') as f:
                    pickle.dump(G_undir, f, protocol=4)
            else:
                np.save(os.path.join(d, name + '.npy'), arr)


def load_dataset(folder):
    """(X, [B1, B2], y, train_mask, test_mask, G_undir, last_nodes, target_nodes) — synthetic_data_gen.py:430-447."""
    import networkx as nx
    with open(os.path.join(folder, 'G_undir.pkl'), 'rb') as f:
        G_undir = pickle.load(f)
    G_undir = nx.relabel_nodes(G_undir, {node: int(node) for node in G_undir.nodes})
    ld = lambda name: np.load(os.path.join(folder, name + '.npy'))
    return ld('flows_in'), [ld('B1'), ld('B2')], ld('targets'), ld('train_mask'), ld('test_mask'), G_undir, \
        ld('last_nodes'), ld('target_nodes')


# ---------------------------------------------------------------------------------------------------------------
# 2. sparse datasets for large complexes
# ---------------------------------------------------------------------------------------------------------------
class SparseDataset:
    """Edge / triangle lists + per-trajectory (edge, sign) lists: the same information as the reference folder."""

    FIELDS = ('n_nodes', 'edges', 'faces', 'traj_ptr', 'flow_edge', 'flow_val', 'last_nodes', 'target_nodes',
              'target_idx', 'train_mask', 'test_mask', 'max_degree')

    def __init__(self, **kw):
        for k in self.FIELDS:
            setattr(self, k, kw[k])

    @property
    def n_traj(self):
        return len(self.last_nodes)

    def save(self, path):
        np.savez_compressed(path, **{k: getattr(self, k) for k in self.FIELDS})

    @classmethod
    def load(cls, path):
        d = np.load(path)
        return cls(**{k: d[k] for k in cls.FIELDS})

    @classmethod
    def from_dense(cls, X, B1, B2, y, train_mask, test_mask, last_nodes, target_nodes):
        from .complex import flows_to_csr, incidence_lists_from_dense
        en, es, te, ts = incidence_lists_from_dense(B1, B2)
        assert np.all(es == np.array([-1, 1])), 'reference orientation expected'
        faces = np.array([sorted(set(en[te[f]].ravel().tolist())) for f in range(len(te))], dtype=np.int32).reshape(-1, 3)
        ptr, fe, fv = flows_to_csr(X)
        return cls(n_nodes=np.int64(B1.shape[0]), edges=en, faces=faces, traj_ptr=ptr, flow_edge=fe, flow_val=fv,
                   last_nodes=np.asarray(last_nodes, np.int32), target_nodes=np.asarray(target_nodes, np.int32),
                   target_idx=np.argmax(np.asarray(y).reshape(len(last_nodes), -1), axis=1).astype(np.int32),
                   train_mask=np.asarray(train_mask, np.int8), test_mask=np.asarray(test_mask, np.int8),
                   max_degree=np.int64(np.asarray(y).shape[1]))

    def to_dense(self):
        """(X [B,E,1], B1, B2, y [B,D,1]) float64 — only for complexes small enough to densify."""
        N, E, F, B, D = int(self.n_nodes), len(self.edges), len(self.faces), self.n_traj, int(self.max_degree)
        X = np.zeros((B, E, 1))
        X[np.repeat(np.arange(B), np.diff(self.traj_ptr)), self.flow_edge, 0] = self.flow_val
        B1 = np.zeros((N, E))
        B1[self.edges[:, 0], np.arange(E)] = -1
        B1[self.edges[:, 1], np.arange(E)] = 1
        lut = {(int(a), int(b)): i for i, (a, b) in enumerate(self.edges)}
        B2 = np.zeros((E, F))
        for j, (a, b, c) in enumerate(self.faces):
            B2[lut[(a, b)], j], B2[lut[(b, c)], j], B2[lut[(a, c)], j] = 1, 1, -1
        y = np.zeros((B, D, 1))
        y[np.arange(B), self.target_idx, 0] = 1
        return X, B1, B2, y


def generate_sparse_dataset(n, m, holes=True, seed=1030, n_waypoints=48, verbose=False, cuts_per_walk=1):
    """Large-complex generator: same complex recipe as random_SC_graph (seeds 1 / 1030), walks BEGIN->A_k->B_k->END
    along BFS shortest paths through `n_waypoints` random waypoints per region (one BFS tree per waypoint instead of
    three nx.shortest_path calls per walk), random truncation as split_paths, 80/20 split.  Deterministic in `seed`.
    cuts_per_walk > 1: every accepted walk is truncated at that many independently drawn points, each a trajectory of its own
    (different prefix, last node and target) — the cheap way to a 32768-trajectory batch."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import breadth_first_order
    coords, valid_idxs, faces, edges = _complex_arrays(n, holes)
    N, E = n, len(edges)
    rng = np.random.RandomState(seed)
    adj = csr_matrix((np.ones(2 * E, np.int8), (np.r_[edges[:, 0], edges[:, 1]], np.r_[edges[:, 1], edges[:, 0]])), shape=(N, N))
    deg = np.diff(adj.indptr)
    D = int(deg.max())
    BEGIN, END, A, B = _regions(coords, valid_idxs)
    ekey = edges[:, 0].astype(np.int64) * N + edges[:, 1]          # sorted (edges are lexicographic)

    def tree(root):
        _, pred = breadth_first_order(adj, int(root), directed=False, return_predecessors=True)
        return pred

    def pick(sel):
        sel = sel[deg[sel] > 0]
        return rng.choice(sel, size=min(n_waypoints, len(sel)), replace=False)
    wp_A, wp_B = [pick(a) for a in A], [pick(b) for b in B]
    wp_END = pick(END)
    trees = {int(r): tree(r) for r in np.concatenate(wp_A + wp_B + [wp_END])}
    begin_ok = BEGIN[deg[BEGIN] > 0]

    def walk_to(src, root):          # path src -> root along root's BFS tree
        pred, out, v = trees[int(root)], [int(src)], int(src)
        while v != int(root):
            v = int(pred[v])
            if v < 0:
                return None
            out.append(v)
        return out

    ptr, fe, fv, last_nodes, target_nodes, target_idx = [0], [], [], [], [], []
    i = 0
    while len(last_nodes) < m:
        k = i % 3
        v0, v1, v2, v3 = rng.choice(begin_ok), rng.choice(wp_A[k]), rng.choice(wp_B[k]), rng.choice(wp_END)
        legs = [walk_to(v0, v1), walk_to(v1, v2), walk_to(v2, v3)]
        if any(l is None for l in legs):
            continue
        path = legs[0][:-1] + legs[1][:-1] + legs[2]
        if len(path) != len(set(path)) or len(path) < 8:
            continue
        full = path
        for _ in range(cuts_per_walk):
            if len(last_nodes) >= m:
                break
            path = full[:4 + rng.choice(range(2, len(full) - 4))]   # split_paths truncation
            prefix, nxt = np.asarray(path[:-2]), path[-2]
            a, b = prefix[:-1], prefix[1:]
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            eid = np.searchsorted(ekey, lo.astype(np.int64) * N + hi)
            order = np.argsort(eid)
            fe.append(eid[order].astype(np.int32))
            fv.append(np.where(a < b, 1.0, -1.0).astype(np.float32)[order])
            ptr.append(ptr[-1] + len(eid))
            last = int(prefix[-1])
            nb = adj.indices[adj.indptr[last]:adj.indptr[last + 1]]
            last_nodes.append(last)
            target_nodes.append(nxt)
            target_idx.append(int(np.searchsorted(np.sort(nb), nxt)))
        i += 1
        if verbose and i % 1000 == 0:
            print('walks:', i)
    train_mask = np.asarray([1] * int(m * 0.8) + [0] * (m - int(m * 0.8)), np.int8)
    rng.shuffle(train_mask)
    return SparseDataset(n_nodes=np.int64(N), edges=edges.astype(np.int32), faces=faces.astype(np.int32),
                         traj_ptr=np.asarray(ptr, np.int32), flow_edge=np.concatenate(fe), flow_val=np.concatenate(fv),
                         last_nodes=np.asarray(last_nodes, np.int32), target_nodes=np.asarray(target_nodes, np.int32),
                         target_idx=np.asarray(target_idx, np.int32), train_mask=train_mask, test_mask=(1 - train_mask).astype(np.int8),
                         max_degree=np.int64(D))


if __name__ == '__main__':
    generate_dataset(400, 1000, 'synthetic')
