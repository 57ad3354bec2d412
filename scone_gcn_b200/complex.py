"""Device-resident simplicial complex (opaque C handle) + host-side conversions.

Replaces the dense operator assembly of the reference (trajectory_experiments.py:239-303): instead of
E x E float64 matrices the complex keeps incidence-derived CSR index arrays on the GPU.  Dense B1 / B2
(the reference's on-disk format, synthetic_data_gen.py:13-14) are accepted and converted exactly by
reading their non-zero patterns, the way the reference derives E_lookup (trajectory_experiments.py:263-268).
"""
import ctypes as C

import numpy as np

from . import _lib

MODEL_IDS = {'scone': 0, 'ebli': 1}


def incidence_lists_from_dense(B1, B2):
    """(edge_nodes [E,2], edge_signs [E,2], tri_edges [F,3], tri_signs [F,3]) from dense B1 [N,E], B2 [E,F]."""
    B1 = np.asarray(B1)
    B2 = np.asarray(B2)
    N, E = B1.shape
    cols, rows = np.nonzero(B1.T)                       # column-major: the two end nodes of edge 0, edge 1, ...
    if len(cols) != 2 * E or not np.array_equal(cols, np.repeat(np.arange(E), 2)):
        raise ValueError('B1 must have exactly two non-zeros per column')
    edge_nodes = rows.reshape(E, 2).astype(np.int32)
    edge_signs = B1[rows, cols].reshape(E, 2)
    if not np.all(np.abs(edge_signs) == 1):
        raise ValueError('B1 entries must be +-1')
    F = B2.shape[1] if B2.ndim == 2 else 0
    if F:
        fcols, frows = np.nonzero(B2.T)
        if len(fcols) != 3 * F or not np.array_equal(fcols, np.repeat(np.arange(F), 3)):
            raise ValueError('B2 must have exactly three non-zeros per column')
        tri_edges = frows.reshape(F, 3).astype(np.int32)
        tri_signs = B2[frows, fcols].reshape(F, 3)
        if not np.all(np.abs(tri_signs) == 1):
            raise ValueError('B2 entries must be +-1')
    else:
        tri_edges = np.zeros((0, 3), np.int32)
        tri_signs = np.zeros((0, 3))
    return edge_nodes, edge_signs.astype(np.int8), tri_edges, tri_signs.astype(np.int8)


def incidence_lists_from_simplices(edges, faces):
    """Same, from sorted edge list [E,2] (a<b) and sorted face list [F,3] (a<b<c), with the reference's sign
    convention (synthetic_data_gen.py:139-161): B1 = (-1 tail, +1 head); B2 = (+ (a,b), + (b,c), - (a,c))."""
    edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
    faces = np.ascontiguousarray(faces, dtype=np.int64).reshape(-1, 3)
    E = len(edges)
    nmax = int(edges.max()) + 1
    keys = edges[:, 0] * nmax + edges[:, 1]
    order = np.argsort(keys, kind='stable')
    skeys = keys[order]

    def eid(a, b):
        k = a * nmax + b
        pos = np.searchsorted(skeys, k)
        if np.any(pos >= E) or np.any(skeys[np.minimum(pos, E - 1)] != k):
            raise ValueError('face side is not an edge of the complex')
        return order[pos]
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
    trip = np.stack([eid(a, b), eid(b, c), eid(a, c)], axis=1)          # signs (+, +, -)
    sgn = np.tile(np.array([1, 1, -1], dtype=np.int8), (len(faces), 1))
    srt = np.argsort(trip, axis=1, kind='stable')                      # store in ascending edge order, like nonzero(B2.T)
    tri_edges = np.take_along_axis(trip, srt, axis=1).astype(np.int32)
    tri_signs = np.take_along_axis(sgn, srt, axis=1).astype(np.int8)
    edge_signs = np.tile(np.array([-1, 1], dtype=np.int8), (E, 1))
    return edges.astype(np.int32), edge_signs, tri_edges, tri_signs


def flows_to_csr(X):
    """Dense flows [B, E, 1] or [B, E] -> (traj_ptr [B+1] int32, flow_edge int32, flow_val float32)."""
    X = np.asarray(X)
    if X.ndim == 3:
        X = X[:, :, 0]
    rows, cols = np.nonzero(X)
    ptr = np.zeros(X.shape[0] + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=X.shape[0]), out=ptr[1:])
    return ptr, cols.astype(np.int32), X[rows, cols].astype(np.float32)


class SimplicialComplex:
    """Owner of a `scone_complex*`.  model in {'scone','ebli'} selects the shift pair."""

    def __init__(self, n_nodes, edge_nodes, edge_signs, tri_edges, tri_signs, model='scone', index_only=False):
        L = _lib.lib()
        self.model = model
        self._edge_nodes = np.ascontiguousarray(edge_nodes, dtype=np.int32)
        self._edge_signs = None if edge_signs is None else np.ascontiguousarray(edge_signs, dtype=np.int8)
        self._tri_edges = np.ascontiguousarray(tri_edges, dtype=np.int32).reshape(-1, 3)
        self._tri_signs = np.ascontiguousarray(tri_signs, dtype=np.int8).reshape(-1, 3)
        h = C.c_void_p()
        self.index_only = bool(index_only)
        create = L.scone_complex_create_index_only if index_only else L.scone_complex_create
        rc = create(int(n_nodes), len(self._edge_nodes), len(self._tri_edges),
                    _lib.ptr(self._edge_nodes), _lib.ptr(self._edge_signs),
                    _lib.ptr(self._tri_edges), _lib.ptr(self._tri_signs), MODEL_IDS[model], C.byref(h))
        _lib.check(rc, 'scone_complex_create')
        self.handle = h
        n, e, f, d = (C.c_int32() for _ in range(4))
        z0, z1 = C.c_int64(), C.c_int64()
        _lib.check(L.scone_complex_dims(h, n, e, f, d, z0, z1), 'scone_complex_dims')
        self.N, self.E, self.F, self.D = n.value, e.value, f.value, d.value
        self.nnz = (z0.value, z1.value)
        self._nbrhoods = None

    @classmethod
    def from_dense(cls, B1, B2, model='scone', flips=None, index_only=False):
        en, es, te, ts = incidence_lists_from_dense(B1, B2)
        if flips is not None:                 # B1 F and F B2 (trajectory_experiments.py:214-219,242-244,290)
            fl = np.asarray(flips).astype(np.int8)
            es = es * fl[:, None]
            ts = ts * fl[te]
        return cls(np.asarray(B1).shape[0], en, es, te, ts, model, index_only=index_only)

    @classmethod
    def from_simplices(cls, n_nodes, edges, faces, model='scone', index_only=False):
        return cls(n_nodes, *incidence_lists_from_simplices(edges, faces), model=model, index_only=index_only)

    def with_model(self, model):
        return self if model == self.model else SimplicialComplex(self.N, self._edge_nodes, self._edge_signs,
                                                                  self._tri_edges, self._tri_signs, model, index_only=self.index_only)

    def shift_csr(self, which):
        rowptr = np.zeros(self.E + 1, np.int32)
        col = np.zeros(self.nnz[which], np.int32)
        val = np.zeros(self.nnz[which], np.float32)
        _lib.check(_lib.lib().scone_complex_get_shift_csr(self.handle, which, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val)))
        return rowptr, col, val

    def shift_dense(self, which):
        rowptr, col, val = self.shift_csr(which)
        M = np.zeros((self.E, self.E))
        M[np.repeat(np.arange(self.E), np.diff(rowptr)), col] = val
        return M

    @property
    def edge_rank(self):
        """rank[e] = internal device row of the caller's edge e (device tensors are stored in a locality order)."""
        a = np.zeros(self.E, np.int32)
        _lib.check(_lib.lib().scone_complex_get_edge_rank(self.handle, _lib.ptr(a)))
        return a

    @property
    def nbrhoods(self):
        if self._nbrhoods is None:
            a = np.zeros((self.N, max(self.D, 1)), np.int32)
            _lib.check(_lib.lib().scone_complex_get_nbrhoods(self.handle, _lib.ptr(a)))
            self._nbrhoods = a[:, :self.D]
        return self._nbrhoods

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().scone_complex_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
