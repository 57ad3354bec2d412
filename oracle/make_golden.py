#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN PYTHON, unmodified, in this container.

TEST INFRASTRUCTURE ONLY.  Run from the repo root in the build container (it reads /root/reference,
which does not exist on the GPU box; the committed .npz files are what travels):

    python oracle/make_golden.py [--skip-default-model]

What runs unchanged from /root/reference/trajectory_analysis:
  synthetic_data_gen.py   random_SC_graph, incidence_matrices, generate_random_walks, path_dataset,
                          generate_dataset, load_dataset                      (NumPy/SciPy/NetworkX: real)
  bunch_model_matrices.py compute_shift_matrices                              (NumPy: real)
  trajectory_experiments.py  hyperparams, data_setup, Bconds_func, scone_func, ebli_func, bunch_func
  scone_trajectory_model.py  Scone_GCN.setup/generate_weights/loss/accuracy/train/test
The last two import `jax`, which is not installable here; they run over oracle/refshim/jax, a
torch-CPU stand-in (fp32 like JAX's default; see oracle/refshim/README.md).  So integer/data
fixtures are produced by the real reference stack, and model fixtures by the reference's model
code with torch doing the fp32 arithmetic that XLA:CPU would do.
"""
import argparse
import importlib
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference/trajectory_analysis'
OUT = os.path.join(REPO, 'tests', 'golden')

sys.path.insert(0, os.path.join(HERE, 'refshim'))
sys.path.insert(1, REF)

import numpy as onp  # noqa: E402
import torch  # noqa: E402
import compat  # noqa: E402

compat.apply()
torch.set_num_threads(8)


def t2n(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return onp.asarray(x)


def import_reference(argv):
    """(Re-)import the reference modules with the given command line (flags are parsed at import:
    trajectory_experiments.py:119)."""
    sys.argv = ['trajectory_experiments.py'] + argv
    for m in ('trajectory_experiments', 'scone_trajectory_model', 'synthetic_data_gen',
              'bunch_model_matrices', 'markov_model'):
        sys.modules.pop(m, None)
    te = importlib.import_module('trajectory_experiments')
    stm = importlib.import_module('scone_trajectory_model')
    sdg = importlib.import_module('synthetic_data_gen')
    return te, stm, sdg


def sparse_dataset(sdg, folder, n):
    """Sparse restatement of what generate_dataset wrote (dense files are too big to commit)."""
    X, (B1, B2), y, train_mask, test_mask, G_undir, last_nodes, target_nodes = sdg.load_dataset(folder)
    G, V, E, faces, edge_to_idx, coords, valid_idxs = sdg.random_SC_graph(n)   # deterministic (seeds 1, 1030)
    rev_X = onp.load(os.path.join(folder, 'rev_flows_in.npy'))
    d = dict(
        n_nodes=onp.int64(n), edges=onp.asarray(E, dtype=onp.int32), faces=onp.asarray(faces, dtype=onp.int32),
        coords=coords, valid_idxs=valid_idxs.astype(onp.int32),
        B1_nz=onp.stack(onp.nonzero(B1)).astype(onp.int32), B1_val=B1[onp.nonzero(B1)].astype(onp.int8),
        B2_nz=onp.stack(onp.nonzero(B2)).astype(onp.int32), B2_val=B2[onp.nonzero(B2)].astype(onp.int8),
        flows_nz=onp.stack(onp.nonzero(X[:, :, 0])).astype(onp.int32),
        flows_val=X[:, :, 0][onp.nonzero(X[:, :, 0])].astype(onp.int8),
        rev_flows_nz=onp.stack(onp.nonzero(rev_X[:, :, 0])).astype(onp.int32),
        rev_flows_val=rev_X[:, :, 0][onp.nonzero(rev_X[:, :, 0])].astype(onp.int8),
        targets_argmax=onp.argmax(y[:, :, 0], axis=1).astype(onp.int32),
        targets_rowsum=y[:, :, 0].sum(axis=1),
        train_mask=train_mask.astype(onp.int8), test_mask=test_mask.astype(onp.int8),
        last_nodes=onp.asarray(last_nodes, dtype=onp.int32), target_nodes=onp.asarray(target_nodes, dtype=onp.int32),
        rev_last_nodes=onp.load(os.path.join(folder, 'rev_last_nodes.npy')).astype(onp.int32),
        rev_target_nodes=onp.load(os.path.join(folder, 'rev_target_nodes.npy')).astype(onp.int32),
        rev_targets_argmax=onp.argmax(onp.load(os.path.join(folder, 'rev_targets.npy'))[:, :, 0], axis=1).astype(onp.int32),
        max_degree=onp.int64(y.shape[1]),
        B1B2_maxabs=onp.float64(onp.abs(B1 @ B2).max()),
    )
    # dense shift operators as the reference builds them (trajectory_experiments.py:240-241,251-253)
    L_lower, L_upper = B1.T @ B1, B2 @ B2.T
    for name, M in (('L_lower', L_lower), ('L_upper', L_upper), ('L1', L_lower + L_upper),
                    ('L1sq', (L_lower + L_upper) @ (L_lower + L_upper))):
        nz = onp.nonzero(M)
        d[name + '_nz'] = onp.stack(nz).astype(onp.int32)
        d[name + '_val'] = M[nz].astype(onp.int32)
    return d


def model_fixture(te, stm, suffix, model, hidden, epochs, batch_size, big_scale, lr=1e-3, wd=5e-5):
    """Run the reference's data_setup / Scone_GCN on dataset `suffix`; return arrays to pin."""
    te.HYPERPARAMS.update({'model': model, 'hidden_layers': hidden, 'flip_edges': 0})
    inputs_all, y_all, train_mask, test_mask, shifts, G_undir, E_lookup, nbrhoods, n_nbrs, tn_all, prefixes = \
        te.data_setup(hops=(1, 2), load=True, folder_suffix=suffix)
    inputs, y_np = inputs_all[0], y_all[0]
    # JAX accepts the NumPy target array implicitly; the torch stand-in needs it converted up front
    import jax.numpy as jnp
    y = jnp.array(y_np)
    in_axes = tuple(([None] * len(shifts)) + [None, None, 0, 0])          # trajectory_experiments.py:325
    func = {'scone': te.scone_func, 'ebli': te.ebli_func, 'bunch': te.bunch_func}[model]

    onp.random.seed(1030)                      # == fresh import of scone_trajectory_model (:15)
    net = stm.Scone_GCN(epochs, lr, batch_size, wd, verbose=False)
    net.setup(func, hidden, shifts, inputs, y, in_axes, train_mask, model_type=model)
    w_init = [onp.array(w) for w in net.weights]

    out = dict(model=model, hidden=onp.asarray(hidden, dtype=onp.int32), lr=lr, wd=wd,
               epochs=epochs, batch_size=batch_size,
               nbrhoods=t2n(nbrhoods).astype(onp.int32), n_nbrs=onp.asarray(n_nbrs, dtype=onp.int32),
               n_weights=len(w_init))
    for i, w in enumerate(w_init):
        out['w_init_%d' % i] = w
    for i, s in enumerate(shifts):
        out['shift_%d' % i] = onp.asarray(s)

    # a deterministic batch mask, independent of the training stream
    rs = onp.random.RandomState(7)
    bm = onp.zeros(len(y), dtype=bool)
    bm[rs.permutation(len(y))[:batch_size]] = True
    bm = onp.logical_and(bm, train_mask)
    out['batch_mask'] = bm

    def evaluate(tag, weights):
        net.weights = weights
        lp = net.model(weights, *shifts, *inputs)
        out[tag + '_logprobs'] = t2n(lp)
        out[tag + '_loss_train'] = t2n(net.loss(weights, inputs, y, train_mask))
        out[tag + '_loss_batch'] = t2n(net.loss(weights, inputs, y, bm))
        g = stm.grad(net.loss)(weights, inputs, y, bm)
        for i, gi in enumerate(g):
            out[tag + '_grad_%d' % i] = t2n(gi)
        out[tag + '_acc_train'] = t2n(net.accuracy(shifts, inputs, y, train_mask, n_nbrs))
        out[tag + '_acc_test'] = t2n(net.accuracy(shifts, inputs, y, test_mask, n_nbrs))

    evaluate('init', [w.copy() for w in w_init])
    rs = onp.random.RandomState(11)
    w_big = [big_scale * rs.randn(*w.shape) for w in w_init]
    for i, w in enumerate(w_big):
        out['w_big_%d' % i] = w
    evaluate('big', w_big)

    if epochs > 0:
        # full reference training loop from the reference's own init + RNG stream
        onp.random.seed(1030)
        net = stm.Scone_GCN(epochs, lr, batch_size, wd, verbose=False)
        net.setup(func, hidden, shifts, inputs, y, in_axes, train_mask, model_type=model)
        tr = net.train(inputs, y, train_mask, test_mask, n_nbrs)
        out['train_result'] = onp.asarray([float(v) for v in tr])
        for i, w in enumerate(net.weights):
            out['w_trained_%d' % i] = t2n(w)
        out['trained_logprobs'] = t2n(net.model(net.weights, *shifts, *inputs))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--skip-default-model', action='store_true')
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix='scone_golden_')
    os.chdir(work)
    te, stm, sdg = import_reference([])
    sdg.color_faces = lambda *a, **k: None          # plotting only (synthetic_data_gen.py:63-80)

    x64 = os.environ.get('REFSHIM_X64', '0') == '1'
    tag = '_f64' if x64 else ''

    # ---- datasets: the reference generator, unchanged -----------------------------------------
    sdg.generate_dataset(400, 1000, 'default', holes=True)       # synthetic_data_gen.py:518-520
    sdg.generate_dataset(120, 60, 'small', holes=True)
    if not x64 and args.only in ('', 'data'):
        onp.savez_compressed(os.path.join(OUT, 'dataset_default.npz'),
                             **sparse_dataset(sdg, 'trajectory_data_1hop_default', 400))
        onp.savez_compressed(os.path.join(OUT, 'dataset_small.npz'),
                             **sparse_dataset(sdg, 'trajectory_data_1hop_small', 120))
        d2 = sparse_dataset(sdg, 'trajectory_data_2hop_small', 120)
        onp.savez_compressed(os.path.join(OUT, 'dataset_small_2hop.npz'),
                             **{k: d2[k] for k in ('flows_nz', 'flows_val', 'targets_argmax', 'last_nodes',
                                                   'target_nodes')})
        print('datasets written')

    # ---- model fixtures on the small complex ----------------------------------------------------
    if args.only in ('', 'small'):
        for model, hidden in (('scone', [(3, 16), (3, 16), (3, 16)]),
                              ('ebli', [(3, 16), (3, 16), (3, 16)]),
                              ('bunch', [(7, 8), (7, 8), (7, 8)]),
                              ('scone', [(3, 32), (3, 32), (3, 32)])):
            fx = model_fixture(te, stm, 'small', model, hidden, epochs=3, batch_size=16, big_scale=0.35)
            name = 'model_small_%s_h%d%s.npz' % (model, hidden[0][1], tag)
            onp.savez_compressed(os.path.join(OUT, name), **fx)
            print('wrote', name, 'train_result', fx.get('train_result'))

    # ---- one forward/grad on the default complex (slow: per-sample dense E x E products) --------
    if not args.skip_default_model and args.only in ('', 'default'):
        fx = model_fixture(te, stm, 'default', 'scone', [(3, 16), (3, 16), (3, 16)], epochs=0, batch_size=100,
                           big_scale=0.3)
        for k in list(fx):
            if k.startswith('shift_'):
                del fx[k]
        onp.savez_compressed(os.path.join(OUT, 'model_default_scone_h16%s.npz' % tag), **fx)
        print('wrote default model fixture; init loss', fx['init_loss_train'])

    # ---- -model bunch on the default complex (BASELINE config 3 on the reference's own default dataset) ----
    if not args.skip_default_model and args.only in ('', 'default_bunch'):
        fx = model_fixture(te, stm, 'default', 'bunch', [(7, 8), (7, 8), (7, 8)], epochs=0, batch_size=100, big_scale=0.3)
        for k in list(fx):
            if k.startswith('shift_'):
                del fx[k]
        onp.savez_compressed(os.path.join(OUT, 'model_default_bunch_h8%s.npz' % tag), **fx)
        print('wrote default bunch fixture; init loss', fx['init_loss_train'])


if __name__ == '__main__':
    main()
