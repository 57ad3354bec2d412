"""Pin the oracle (oracle/scone_oracle.py) against golden vectors produced by the reference's own
Python (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from golden_util import Dataset, load, weights_of
from oracle import scone_oracle as so


@pytest.fixture(scope='module', params=['dataset_default.npz', 'dataset_small.npz'])
def ds(request):
    return Dataset(request.param)


def test_survey_facts_default():
    d = Dataset('dataset_default.npz')
    assert (d.N, d.E, d.F, d.D) == (400, 1001, 649, 13)
    assert d.edges[:5].tolist() == [[0, 1], [0, 2], [0, 4], [0, 19], [1, 2]]
    assert d.faces[:3].tolist() == [[0, 1, 2], [0, 1, 19], [0, 2, 4]]
    assert d.n_traj == 1000 and d.train_mask.sum() == 800 and d.test_mask.sum() == 200
    assert float(d.raw['B1B2_maxabs']) == 0.0


def test_incidence_bit_exact(ds):
    B1, B2 = so.incidence_matrices(ds.N, ds.edges, ds.faces)
    assert np.array_equal(B1, ds.B1) and np.array_equal(B2, ds.B2)
    # invariants: 2 nnz per B1 column {-1,+1}; 3 per B2 column, (+,-,+) in ascending edge order
    assert np.all((B1 != 0).sum(0) == 2) and np.all(B1.sum(0) == 0)
    for f in range(ds.F):
        nz = np.nonzero(B2[:, f])[0]
        assert B2[nz, f].tolist() == [1.0, -1.0, 1.0]
    assert np.abs(B1 @ B2).max() == 0


def test_shift_matrices_bit_exact(ds):
    Ll, Lu = so.shift_matrices(ds.B1, ds.B2, 'scone')
    assert np.array_equal(Ll, ds.shift('L_lower')) and np.array_equal(Lu, ds.shift('L_upper'))
    L1, L1sq = so.shift_matrices(ds.B1, ds.B2, 'ebli')
    assert np.array_equal(L1, ds.shift('L1')) and np.array_equal(L1sq, ds.shift('L1sq'))


def test_flows_targets_bit_exact(ds):
    lut = {(int(a), int(b)): i for i, (a, b) in enumerate(ds.edges)}
    nbrs = so.adjacency_from_B1(ds.B1)
    # rebuild each prefix path from its flow (simple paths) and re-encode it
    for t in range(0, ds.n_traj, 7):
        f = ds.flows[t, :, 0]
        succ = {}
        for e in np.nonzero(f)[0]:
            a, b = ds.edges[e]
            (u, v) = (a, b) if f[e] > 0 else (b, a)
            succ[int(u)] = int(v)
        start = (set(succ) - set(succ.values())).pop()
        path = [start]
        while path[-1] in succ:
            path.append(succ[path[-1]])
        assert path[-1] == ds.last_nodes[t]
        assert np.array_equal(so.path_to_flow(path, lut, ds.E), ds.flows[t])
        oh = so.neighborhood_to_onehot(np.array(nbrs[path[-1]]), ds.target_nodes[t], ds.D)
        assert np.array_equal(oh, ds.targets[t])


@pytest.mark.parametrize('name', ['model_small_scone_h16', 'model_small_ebli_h16', 'model_small_bunch_h8',
                                  'model_small_scone_h32'])
def test_model_forward_loss_grad_vs_reference(name):
    ds = Dataset('dataset_small.npz')
    for suffix, dt, rtol in (('', torch.float32, 2e-5), ('_f64', torch.float64, 1e-10)):
        fx = load(name + suffix + '.npz')
        model = str(fx['model'])
        shifts = so.shift_matrices(ds.B1, ds.B2, model)
        for i, s in enumerate(shifts):
            assert np.allclose(s, fx['shift_%d' % i], rtol=0, atol=1e-12)
        nbrhoods, n_nbrs, _ = so.neighbourhood_tables(ds.B1, ds.last_nodes)
        assert np.array_equal(nbrhoods, fx['nbrhoods']) and np.array_equal(n_nbrs, fx['n_nbrs'])
        orc = so.DenseOracle(model, shifts, ds.B1, ds.last_nodes, ds.flows, ds.targets, dtype=dt)
        hidden = [tuple(h) for h in fx['hidden'].tolist()]
        w0 = so.generate_weights(np.random.RandomState(1030), 1, hidden, 1, model)
        for a, b in zip(w0, weights_of(fx, 'w_init')):
            assert np.array_equal(a, b)
        wd = float(fx['wd'])
        for tag in ('init', 'big'):
            W = weights_of(fx, 'w_' + tag)
            with torch.no_grad():
                lp = orc.forward(W).numpy()
            ref = fx[tag + '_logprobs']
            assert np.allclose(lp, ref, rtol=rtol, atol=rtol), np.abs(lp - ref).max()
            l, g = orc.loss_and_grads(W, fx['batch_mask'], wd)
            assert np.allclose(l, fx[tag + '_loss_batch'], rtol=rtol)
            for i, gi in enumerate(g):
                r = fx['%s_grad_%d' % (tag, i)]
                assert np.abs(gi - r).max() <= 10 * rtol * max(np.abs(r).max(), 1e-30), (tag, i)
            assert orc.accuracy(W, ds.train_mask) == pytest.approx(float(fx[tag + '_acc_train']), abs=1e-7)
            assert orc.accuracy(W, ds.test_mask) == pytest.approx(float(fx[tag + '_acc_test']), abs=1e-7)


@pytest.mark.parametrize('name', ['model_small_scone_h16_f64', 'model_small_ebli_h16_f64'])
def test_training_loop_vs_reference(name):
    """Same init, same batch-mask RNG stream, same Adam -> same weights after 9 steps (fp64 twin)."""
    ds = Dataset('dataset_small.npz')
    fx = load(name + '.npz')
    model = str(fx['model'])
    orc = so.DenseOracle(model, so.shift_matrices(ds.B1, ds.B2, model), ds.B1, ds.last_nodes, ds.flows, ds.targets,
                         dtype=torch.float64)
    rng = np.random.RandomState(1030)
    hidden = [tuple(h) for h in fx['hidden'].tolist()]
    W0 = so.generate_weights(rng, 1, hidden, 1, model)
    W, res = so.train(orc, rng, W0, ds.train_mask, ds.test_mask, int(fx['epochs']), int(fx['batch_size']),
                      float(fx['lr']), float(fx['wd']), dtype=np.float64)
    for i, w in enumerate(W):
        assert np.allclose(w, fx['w_trained_%d' % i], rtol=1e-8, atol=1e-12), i
    assert np.allclose(res, fx['train_result'], rtol=1e-8)


def test_default_complex_forward_vs_reference():
    ds = Dataset('dataset_default.npz')
    fx = load('model_default_scone_h16.npz')
    orc = so.DenseOracle('scone', so.shift_matrices(ds.B1, ds.B2, 'scone'), ds.B1, ds.last_nodes, ds.flows, ds.targets)
    assert float(fx['init_loss_train']) == pytest.approx(np.log(13), abs=1e-3)      # Q1: padded zeros in the normaliser
    W = weights_of(fx, 'w_big')
    idx = torch.arange(0, 1000, 9)
    with torch.no_grad():
        lp = orc.forward(W, idx).numpy()
    assert np.allclose(lp, fx['big_logprobs'][::9], rtol=2e-5, atol=2e-5)


def test_default_complex_bunch_forward_vs_reference():
    """-model bunch on the reference's default dataset (400 nodes): the oracle against the reference's own run."""
    ds = Dataset('dataset_default.npz')
    fx = load('model_default_bunch_h8.npz')
    orc = so.DenseOracle('bunch', so.shift_matrices(ds.B1, ds.B2, 'bunch'), ds.B1, ds.last_nodes, ds.flows, ds.targets)
    W = weights_of(fx, 'w_big')
    idx = torch.arange(0, 1000, 9)
    with torch.no_grad():
        lp = orc.forward(W, idx).numpy()
    assert np.allclose(lp, fx['big_logprobs'][::9], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('model', ['scone', 'ebli'])
def test_sparse_oracle_matches_dense(model):
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_%s_h16.npz' % model)
    tri_edges = np.stack([np.nonzero(ds.B2[:, f])[0] for f in range(ds.F)])
    tri_signs = np.stack([ds.B2[tri_edges[f], f] for f in range(ds.F)])
    spo = so.SparseOracle(model, ds.edges, tri_edges, tri_signs, ds.N, dtype=np.float64)
    orc = so.DenseOracle(model, so.shift_matrices(ds.B1, ds.B2, model), ds.B1, ds.last_nodes, ds.flows, ds.targets,
                         dtype=torch.float64)
    W = weights_of(fx, 'w_big')
    X = ds.flows[:, :, 0].T.copy()
    lp = spo.forward(W, X, ds.last_nodes)
    with torch.no_grad():
        ref = orc.forward(W).numpy()[:, :, 0]
    assert np.allclose(lp, ref, rtol=1e-10, atol=1e-12)
    mask = fx['batch_mask'].astype(np.float64)
    nll, g = spo.loss_and_grads(W, X, ds.last_nodes, ds.raw['targets_argmax'], mask, n_total=mask.sum())
    l, gref = orc.loss_and_grads(W, fx['batch_mask'], 0.0)
    assert np.allclose(nll / mask.sum(), l, rtol=1e-10)
    for a, b in zip(g, gref):
        assert np.allclose(a, b, rtol=1e-8, atol=1e-14)


def test_two_target_accuracy_matches_the_reference_run():
    """oracle.two_target_accuracy against the reference's own method (oracle/make_golden_two_target.py: train mask, then test mask,
    global RNG seeded 4242): scores, the redrawn random targets and the RNG stream position afterwards."""
    ds = Dataset('dataset_small.npz')
    fx, tt = load('model_small_scone_h16.npz'), load('two_target_small_scone_h16.npz')
    preds = fx['big_logprobs']
    n_nbrs = fx['n_nbrs']
    np.random.seed(int(tt['seed']))
    tr, rt = so.two_target_accuracy(preds, ds.targets, ds.train_mask, n_nbrs)
    assert tr == float(tt['train']) and np.array_equal(rt, tt['random_targets_after_train'])
    ts, rt = so.two_target_accuracy(preds, ds.targets, ds.test_mask, n_nbrs, random_targets=rt)
    assert ts == float(tt['test']) and np.array_equal(rt, tt['random_targets_after_test'])
    assert np.random.randint(0, 1 << 30) == int(tt['next_draw'])
