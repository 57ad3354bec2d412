"""Golden vectors for Scone_GCN.two_target_accuracy (scone_trajectory_model.py:73-108): runs the reference's own method, unmodified,
on the `small` dataset with the `w_big` weights of tests/golden/model_small_scone_h16.npz, for the train and the test mask in the order
trajectory_experiments.py:490-491 calls them, from a fresh RNG seed.  jax semantics the torch stand-in lacks are supplied here:
`pred_choice[i]` with i past the end CLAMPS to the last element (jax gather), it does not raise.

    python oracle/make_golden_two_target.py        (build container only: needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as onp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                                            # noqa: E402


class ClampVec(torch.Tensor):
    def __getitem__(self, i):
        if isinstance(i, (int, onp.integer)):
            n = self.shape[0]
            i = min(max(int(i), -n), n - 1)
        return torch.Tensor.__getitem__(self.as_subclass(torch.Tensor), i)


def main():
    work = tempfile.mkdtemp(prefix='scone_golden_tt_')
    os.chdir(work)
    te, stm, sdg = mg.import_reference([])
    sdg.color_faces = lambda *a, **k: None
    sdg.generate_dataset(400, 1000, 'default', holes=True)                          # same generator call sequence as make_golden.py
    sdg.generate_dataset(120, 60, 'small', holes=True)
    import jax.numpy as jnp
    real_argmax = jnp.argmax

    def clamping_argmax(x, axis=None):
        r = real_argmax(x, axis)
        return r.as_subclass(ClampVec) if isinstance(r, torch.Tensor) and r.dim() >= 1 else r
    stm.np.argmax = clamping_argmax
    hidden = [(3, 16), (3, 16), (3, 16)]
    te.HYPERPARAMS.update({'model': 'scone', 'hidden_layers': hidden, 'flip_edges': 0})
    inputs_all, y_all, train_mask, test_mask, shifts, G_undir, E_lookup, nbrhoods, n_nbrs, tn_all, prefixes = \
        te.data_setup(hops=(1, 2), load=True, folder_suffix='small')
    inputs, y = inputs_all[0], jnp.array(y_all[0])
    in_axes = tuple(([None] * len(shifts)) + [None, None, 0, 0])
    onp.random.seed(1030)
    net = stm.Scone_GCN(0, 1e-3, 16, 5e-5, verbose=False)
    net.setup(te.scone_func, hidden, shifts, inputs, y, in_axes, train_mask, model_type='scone')
    fx = onp.load(os.path.join(mg.OUT, 'model_small_scone_h16.npz'))
    net.weights = [fx['w_big_%d' % i] for i in range(int(fx['n_weights']))]
    onp.random.seed(4242)
    y_np = onp.asarray(y_all[0])
    tr = net.two_target_accuracy(shifts, inputs, y_np, onp.asarray(train_mask), n_nbrs)
    rt_after_train = onp.array(net.random_targets)
    ts = net.two_target_accuracy(shifts, inputs, y_np, onp.asarray(test_mask), n_nbrs)
    out = dict(seed=4242, train=float(tr), test=float(ts), random_targets_after_train=rt_after_train,
               random_targets_after_test=onp.array(net.random_targets), next_draw=onp.random.randint(0, 1 << 30))
    onp.savez_compressed(os.path.join(mg.OUT, 'two_target_small_scone_h16.npz'), **out)
    print('two-target accuracies', tr, ts, 'targets', rt_after_train[:10])


if __name__ == '__main__':
    main()
