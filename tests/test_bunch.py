"""-model bunch (SCCONV): operator construction on the CPU, CUDA forward/backward parity on the GPU."""
import numpy as np
import pytest

from golden_util import Dataset, load, weights_of
from scone_gcn_b200.bunch_model_matrices import compute_shift_matrices


def test_shift_matrices_match_reference_fixture():
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_bunch_h8.npz')
    S = compute_shift_matrices(ds.B1, ds.B2)
    for k in range(7):
        assert np.allclose(S[k].toarray(), fx['shift_%d' % k], rtol=1e-12, atol=1e-12), k


@pytest.mark.gpu
@pytest.mark.parametrize('which', ['small', 'default'])
def test_bunch_forward_grads_accuracy_vs_reference_golden(which):
    """small: the 120-node complex; default: the reference's own default dataset (400 nodes, 1000 trajectories; BASELINE config 3)."""
    import scone_gcn_b200 as sg
    from scone_gcn_b200.bunch import BunchModel, CsrOperator
    ds = Dataset('dataset_%s.npz' % which)
    fx = load('model_%s_bunch_h8.npz' % which)
    ops = [CsrOperator(M) for M in compute_shift_matrices(ds.B1, ds.B2)]
    net = BunchModel(ops, fx['nbrhoods'], [int(h[1]) for h in fx['hidden']], micro_batch=32)
    ptr, fe, fv = sg.flows_to_csr(ds.flows)
    tgt = ds.raw['targets_argmax']
    mask = fx['batch_mask'].astype(np.float32)
    wd = float(fx['wd'])
    for tag in ('init', 'big'):
        W = weights_of(fx, 'w_' + tag)
        net.set_weights(W)
        lp = net.forward(ptr, fe, fv, ds.last_nodes)
        ref = fx[tag + '_logprobs'][:, :, 0]
        assert np.abs(lp - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (tag, np.abs(lp - ref).max())
        buf = net.loss_grad(ptr, fe, fv, ds.last_nodes, tgt, mask)
        n = net.n_params
        assert buf[n + 1] == mask.sum()
        ridge = wd * sum(float((np.asarray(w, np.float64) ** 2).sum()) for w in W)
        assert buf[n] / buf[n + 1] + ridge == pytest.approx(float(fx[tag + '_loss_batch']), rel=2e-5)
        grads = net.unflatten(buf[:n] / buf[n + 1])
        for i, g in enumerate(grads):
            g = g + 2 * wd * np.asarray(W[i], np.float32)
            r = fx['%s_grad_%d' % (tag, i)]
            assert np.abs(g - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-30), (tag, i)


@pytest.mark.gpu
def test_bunch_training_through_scone_gcn():
    import scone_gcn_b200 as sg
    from scone_gcn_b200.bunch import CsrOperator
    from scone_gcn_b200.scone_trajectory_model import Scone_GCN
    from scone_gcn_b200 import trajectory_experiments as te
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_bunch_h8.npz')
    np.random.seed(1030)
    ops = [CsrOperator(M) for M in compute_shift_matrices(ds.B1, ds.B2)]
    inputs = [fx['nbrhoods'], ds.last_nodes, ds.flows]
    net = Scone_GCN(int(fx['epochs']), float(fx['lr']), int(fx['batch_size']), float(fx['wd']), verbose=False)
    net.setup(te.bunch_func, [(7, 8)] * 3, ops, inputs, ds.targets, None, ds.train_mask, model_type='bunch')
    for a, r in zip(net.weights, weights_of(fx, 'w_init')):
        assert np.array_equal(np.asarray(a), r)
    res = net.train(inputs, ds.targets, ds.train_mask, ds.test_mask, fx['n_nbrs'])
    ref = fx['train_result']
    assert res[0] == pytest.approx(ref[0], rel=1e-4) and res[2] == pytest.approx(ref[2], rel=1e-4)
    assert res[1] == pytest.approx(ref[1], abs=1e-7) and res[3] == pytest.approx(ref[3], abs=1e-7)
    lp1 = te.bunch_func(net.weights, *ops, fx['nbrhoods'], ds.last_nodes[3], ds.flows[3])
    assert lp1.shape == (ds.D, 1) and np.isfinite(lp1).all()
