"""Index construction on hand-built complexes: the library's exact integer CSR operators and neighbour table
against the oracle's dense restatement of the reference (`incidence_matrices`, `B1.T @ B1`, `B2 @ B2.T`, `nbrhoods`) — bit-exact,
no GPU needed (`scone_complex_create_index_only`)."""
import numpy as np
import pytest

from oracle import scone_oracle as so

CASES = {
    # name: (n_nodes, edges (a < b, lexicographic), faces (a < b < c))
    'single_triangle': (3, [(0, 1), (0, 2), (1, 2)], [(0, 1, 2)]),
    'two_triangles_sharing_an_edge': (4, [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3)], [(0, 1, 2), (1, 2, 3)]),
    'path_without_triangles': (5, [(0, 1), (1, 2), (2, 3), (3, 4)], []),
    'isolated_nodes': (6, [(0, 1), (0, 2), (1, 2)], [(0, 1, 2)]),                       # nodes 3, 4, 5 have no edges
    'star_max_degree_and_leaves': (7, [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (1, 2)], [(0, 1, 2)]),
    'triangle_fan_with_open_edge': (5, [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (2, 3)], [(0, 1, 2), (0, 2, 3)]),
}


@pytest.mark.parametrize('model', ['scone', 'ebli'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_shift_operators_and_neighbour_table_bit_exact(name, model):
    import scone_gcn_b200 as sg
    n, edges, faces = CASES[name]
    edges, faces = np.asarray(edges, np.int64).reshape(-1, 2), np.asarray(faces, np.int64).reshape(-1, 3)
    B1, B2 = so.incidence_matrices(n, edges, faces)
    if len(faces):
        assert np.abs(B1 @ B2).max() == 0                                              # boundary of a boundary
    cx = sg.SimplicialComplex.from_simplices(n, edges, faces, model, index_only=True)
    assert (cx.N, cx.E, cx.F) == (n, len(edges), len(faces))
    ref = so.shift_matrices(B1, B2, model)
    for k in range(2):
        got = cx.shift_dense(k)
        assert np.array_equal(got, np.asarray(ref[k])), (name, model, k)
        rowptr, col, _ = cx.shift_csr(k)
        for e in range(cx.E):                                                          # columns ascending: the fixed summation order
            c = col[rowptr[e]:rowptr[e + 1]]
            assert np.all(np.diff(c) > 0)
    # neighbour table: sorted neighbours, padded with -1 to the maximum degree; isolated nodes are all -1
    last_nodes = np.arange(n)
    nb, n_nbrs, _ = so.neighbourhood_tables(B1, last_nodes)
    assert cx.D == nb.shape[1] == int(np.abs(B1).sum(axis=1).max())
    assert np.array_equal(cx.nbrhoods, nb)
    assert np.array_equal((cx.nbrhoods >= 0).sum(axis=1), n_nbrs)
    # the two routes into the library agree (dense B1 / B2 as the reference stores them vs simplex lists)
    cx2 = sg.SimplicialComplex.from_dense(B1, B2, model, index_only=True)
    for k in range(2):
        assert np.array_equal(cx2.shift_dense(k), cx.shift_dense(k))
    assert sorted(np.asarray(cx.edge_rank).tolist()) == list(range(cx.E))              # internal order: a permutation
