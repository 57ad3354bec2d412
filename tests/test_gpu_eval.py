"""Evaluation on the device: loss, next-node accuracy, two-target accuracy and the reverse / regional experiments of the
reference run with the log-probs staying on the GPU (scone_model_eval_host / scone_model_two_target_host); exact integer counts,
golden vectors produced by the reference's own methods."""
import numpy as np
import pytest
import torch

from golden_util import Dataset, load, weights_of
from oracle import scone_oracle as so

pytestmark = pytest.mark.gpu


def _net(ds, model='scone', hidden=((3, 16),) * 3):
    import scone_gcn_b200 as sg
    from scone_gcn_b200.scone_trajectory_model import Scone_GCN
    from scone_gcn_b200 import trajectory_experiments as te
    cx = sg.SimplicialComplex.from_dense(ds.B1, ds.B2, model)
    inputs = [te.Bconds(cx), ds.last_nodes, ds.flows]
    net = Scone_GCN(0, 1e-3, 16, 5e-5, verbose=False)
    net.setup(te.scone_func if model == 'scone' else te.ebli_func, list(hidden), te.shift_handles(cx), inputs, ds.targets, None, ds.train_mask,
              model_type=model)
    return cx, net, inputs, te


def test_two_target_accuracy_matches_the_reference_run():
    """Scone_GCN.two_target_accuracy (device ranking + comparison, host redraw loop) against the reference's own method
    (oracle/make_golden_two_target.py): train mask then test mask as trajectory_experiments.py:490-491, global RNG seeded 4242."""
    ds = Dataset('dataset_small.npz')
    fx, tt = load('model_small_scone_h16.npz'), load('two_target_small_scone_h16.npz')
    cx, net, inputs, te = _net(ds)
    net.weights = weights_of(fx, 'w_big')
    n_nbrs = fx['n_nbrs']
    np.random.seed(int(tt['seed']))
    tr = net.two_target_accuracy(net.shifts, inputs, ds.targets, ds.train_mask, n_nbrs)
    assert tr == float(tt['train']) and np.array_equal(net.random_targets, tt['random_targets_after_train'])
    ts = net.two_target_accuracy(net.shifts, inputs, ds.targets, ds.test_mask, n_nbrs)
    assert ts == float(tt['test']) and np.array_equal(net.random_targets, tt['random_targets_after_test'])
    assert np.random.randint(0, 1 << 30) == int(tt['next_draw'])                     # the host RNG stream is where the reference leaves it


def test_device_loss_accuracy_and_predictions_match_the_reference_fixture():
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_scone_h16.npz')
    cx, net, inputs, te = _net(ds)
    n_nbrs = fx['n_nbrs']
    for tag in ('init', 'big'):
        net.weights = weights_of(fx, 'w_' + tag)
        assert abs(float(net.loss(net.weights, inputs, ds.targets, ds.train_mask)) - float(fx[tag + '_loss_train'])) <= 1e-5 * abs(float(fx[tag + '_loss_train']))
        assert abs(float(net.loss(net.weights, inputs, ds.targets, fx['batch_mask'])) - float(fx[tag + '_loss_batch'])) <= 1e-5 * abs(float(fx[tag + '_loss_batch']))
        assert net.accuracy(net.shifts, inputs, ds.targets, ds.train_mask, n_nbrs) == pytest.approx(float(fx[tag + '_acc_train']), abs=1e-7)         # (the fixture stores float32)
        assert net.accuracy(net.shifts, inputs, ds.targets, ds.test_mask, n_nbrs) == pytest.approx(float(fx[tag + '_acc_test']), abs=1e-7)
    # predictions: argmax with the -100 masking, first maximum, against NumPy on the reference's log-probs (ties included: at the
    # 0.01-scale init every logit of a row is ~0)
    p = net._prepared(inputs)
    net._push(net.weights)
    choice = net._net.evaluate(p.ptr, p.edge, p.val, p.last_nodes, n_nbrs=n_nbrs, want_choice=True)['choice']
    ref = np.array(fx['big_logprobs'][:, :, 0])
    for i in range(len(ref)):
        ref[i, n_nbrs[i]:] = -100
    margin = np.sort(ref, axis=1)[:, -1] - np.sort(ref, axis=1)[:, -2]
    decided = margin > 1e-5
    assert decided.mean() > 0.9 and np.array_equal(choice[decided], np.argmax(ref, axis=1)[decided])


def test_reverse_and_regional_experiments_on_the_device():
    """trajectory_experiments.py:449-453 (regional masks) and :497-504 (reversed test flows): Scone_GCN.test with the reversed inputs /
    regional masks against the dense oracle — loss 1e-5, identical accuracy."""
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_scone_h16.npz')
    cx, net, inputs, te = _net(ds)
    W = weights_of(fx, 'w_big')
    net.weights = W
    raw = ds.raw
    rev_last = raw['rev_last_nodes'].astype(np.int64)
    rev_targets = np.zeros_like(ds.targets)
    rev_targets[np.arange(ds.n_traj), raw['rev_targets_argmax'], 0] = 1.0
    adj = so.adjacency_from_B1(ds.B1)
    rev_n_nbrs = np.array([len(adj[n]) for n in rev_last])
    net.verbose = False
    loss, acc = net.test([inputs[0], rev_last, ds.rev_flows], rev_targets, ds.test_mask, rev_n_nbrs)
    orc = so.DenseOracle('scone', so.shift_matrices(ds.B1, ds.B2, 'scone'), ds.B1, rev_last, ds.rev_flows, rev_targets)
    with torch.no_grad():
        l_ref = float(orc.loss(W, ds.test_mask, 5e-5))
    assert abs(float(loss) - l_ref) <= 1e-5 * abs(l_ref)
    assert acc == pytest.approx(orc.accuracy(W, ds.test_mask), abs=1e-12)
    # regional masks: i % 3 == 1 trains, i % 3 == 2 tests (0: middle, 1: top, 2: bottom)
    reg_test = np.array([1 if i % 3 == 2 else 0 for i in range(ds.n_traj)])
    n_nbrs = fx['n_nbrs']
    loss, acc = net.test(inputs, ds.targets, reg_test, n_nbrs)
    orc = so.DenseOracle('scone', so.shift_matrices(ds.B1, ds.B2, 'scone'), ds.B1, ds.last_nodes, ds.flows, ds.targets)
    with torch.no_grad():
        l_ref = float(orc.loss(W, reg_test, 5e-5))
    assert abs(float(loss) - l_ref) <= 1e-5 * abs(l_ref)
    assert acc == pytest.approx(orc.accuracy(W, reg_test), abs=1e-12)
