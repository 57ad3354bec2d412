"""Host-side mirror of the reference experiment driver (trajectory_analysis/trajectory_experiments.py).

Same flags (hyperparams, :78-117), model functions (scone_func / ebli_func / bunch_func, :137-203),
data_setup return tuple (:206-311) and train_model flow (:313-513).  The dense E x E shift matrices are
replaced by ShiftHandle objects that carry the device-resident complex (they still convert to the dense
matrix on request, for small complexes), and Bconds_func by a callable object doing the same row gather.

    python -m scone_gcn_b200.trajectory_experiments -data_folder_suffix synthetic -model scone -epochs 5
"""
import os
import sys

import numpy as onp

from .complex import SimplicialComplex
from .model import SconeModel
from .bunch import BunchModel, CsrOperator
from .bunch_model_matrices import compute_shift_matrices
from .scone_trajectory_model import Scone_GCN
from . import synthetic_data_gen as sdg

DEFAULTS = {'model': 'scone',
            'epochs': 1000,
            'learning_rate': 0.001,
            'weight_decay': 0.00005,
            'batch_size': 100,
            'hidden_layers': [(3, 16), (3, 16), (3, 16)],
            'describe': 1,
            'reverse': 0,
            'load_data': 1,
            'load_model': 0,
            'markov': 0,
            'model_name': 'model',
            'regional': 0,
            'flip_edges': 0,
            'data_folder_suffix': 'working',
            'multi_graph': '',
            'holes': 1}


def hyperparams(args=None):
    """
    Parse hyperparameters from command line                               (trajectory_experiments.py:78-117)

    For hidden_layers, input [(3, 8), (3, 8)] as 3_8_3_8.  Every non-string flag becomes a float (Q6).
    """
    args = sys.argv if args is None else args
    hp = dict(DEFAULTS)
    hp['hidden_layers'] = list(DEFAULTS['hidden_layers'])
    for i in range(len(args) - 1):
        if args[i][0] == '-':
            if args[i][1:] == 'hidden_layers':
                nums = list(map(int, args[i + 1].split("_")))
                hp['hidden_layers'] = []
                for j in range(0, len(nums), 2):
                    hp['hidden_layers'] += [(nums[j], nums[j + 1])]
            elif args[i][1:] in ['model_name', 'data_folder_suffix', 'multi_graph', 'model']:
                hp[args[i][1:]] = str(args[i + 1])
            else:
                hp[args[i][1:]] = float(args[i + 1])
    return hp


def _is_cli():
    main = sys.modules.get('__main__')
    name = getattr(getattr(main, '__spec__', None), 'name', '') or os.path.basename(getattr(main, '__file__', '') or '')
    return 'trajectory_experiments' in name


# the reference parses sys.argv at import (:119); doing that under pytest / a host application would choke on
# foreign flags, so only the CLI entry point does it.
HYPERPARAMS = hyperparams(sys.argv if _is_cli() else [])


class ShiftHandle:
    """Stands where the reference passes a dense E x E shift matrix; carries the device-resident complex."""

    def __init__(self, cx, which):
        self.complex, self.which = cx, which
        self.shape = (cx.E, cx.E)

    def toarray(self):
        return self.complex.shift_dense(self.which)

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a if dtype is None else a.astype(dtype)


def shift_handles(cx):
    return [ShiftHandle(cx, 0), ShiftHandle(cx, 1)]


class Bconds:
    """Bconds_func (trajectory_experiments.py:298-303): rows of B1 (plus an appended zero row for the -1 padding)
    for the neighbours of node n."""

    def __init__(self, cx, B1=None):
        self.complex = cx
        self._B1 = B1

    def __call__(self, n):
        if self._B1 is None:
            B1 = onp.zeros((self.complex.N + 1, self.complex.E))
            en, es = self.complex._edge_nodes, self.complex._edge_signs
            if es is None:
                es = onp.tile(onp.array([-1, 1], onp.int8), (self.complex.E, 1))
            B1[en[:, 0], onp.arange(self.complex.E)] = es[:, 0]
            B1[en[:, 1], onp.arange(self.complex.E)] = es[:, 1]
            self._B1 = B1
        return self._B1[self.complex.nbrhoods[n]]


_MODEL_CACHE = {}


def _single(model_type, weights, S0, last_node, flow):
    cx = getattr(S0, 'complex', None)
    if cx is None:
        raise TypeError('the shift arguments must be the ShiftHandle objects returned by data_setup / shift_handles')
    n_layers = (len(weights) - 1) / 3
    assert n_layers % 1 == 0, 'wrong number of weights'
    cx = cx.with_model(model_type) if cx.model != model_type else cx
    hidden = tuple(int(onp.asarray(weights[3 * i]).shape[1]) for i in range(int(n_layers)))
    key = (id(cx), hidden)
    net = _MODEL_CACHE.get(key)
    if net is None:
        net = _MODEL_CACHE[key] = SconeModel(cx, hidden, micro_batch=1)
    net.set_weights([onp.asarray(w) for w in weights], reset_adam=False)
    f = onp.asarray(flow).reshape(1, -1)
    from .complex import flows_to_csr
    ptr, fe, fv = flows_to_csr(f)
    return net.forward(ptr, fe, fv, onp.asarray([int(last_node)], onp.int32)).reshape(-1, 1)


def scone_func(weights, S_lower, S_upper, Bcond_func, last_node, flow):
    """
    Forward pass of the SCoNe model with variable number of layers       (trajectory_experiments.py:137-152)
    """
    return _single('scone', weights, S_lower, last_node, flow)


def ebli_func(weights, S_lower, S_upper, Bcond_func, last_node, flow):
    """
    Forward pass of the Ebli model with variable number of layers         (trajectory_experiments.py:155-170)
    """
    return _single('ebli', weights, S_lower, last_node, flow)


def bunch_func(weights, S_00, S_10, S_01, S_11, S_21, S_12, S_22, nbrhoods, last_node, flow):
    """
    Forward pass of the Bunch model                                       (trajectory_experiments.py:173-203)
    """
    assert len(weights) % 7 == 0, 'wrong number of weights'
    shifts = (S_00, S_10, S_01, S_11, S_21, S_12, S_22)
    if not all(hasattr(s, 'handle') for s in shifts):
        raise TypeError('the shift arguments must be the CsrOperator objects returned by data_setup')
    hidden = tuple(int(onp.asarray(weights[7 * i]).shape[1]) for i in range(len(weights) // 7 - 1))
    key = (tuple(id(s) for s in shifts), hidden)
    net = _MODEL_CACHE.get(key)
    if net is None:
        net = _MODEL_CACHE[key] = BunchModel(shifts, onp.asarray(nbrhoods), hidden, micro_batch=1)
    net.set_weights([onp.asarray(w) for w in weights], reset_adam=False)
    from .complex import flows_to_csr
    ptr, fe, fv = flows_to_csr(onp.asarray(flow).reshape(1, -1))
    return net.forward(ptr, fe, fv, onp.asarray([int(last_node)], onp.int32)).reshape(-1, 1)


def data_setup(hops=(1,), load=True, folder_suffix='schaub'):
    """
    Imports and sets up flow, target, and shift handles for model training (trajectory_experiments.py:206-311).
    Returns the same 11-tuple as the reference.
    """
    inputs_all, y_all, target_nodes_all = [], [], []
    flips = None
    if HYPERPARAMS['flip_edges']:
        onp.random.seed(1)
        G_undir = sdg.load_dataset('trajectory_data_1hop_' + folder_suffix)[5]
        flips = onp.random.choice([1, -1], size=len(G_undir.edges), replace=True, p=[0.8, 0.2])
    if not load:
        sdg.generate_dataset(400, 1000, folder=folder_suffix, holes=HYPERPARAMS['holes'])
        raise Exception('Data generation done')
    for h in hops:
        folder = 'trajectory_data_' + str(h) + 'hop_' + folder_suffix
        X, B_matrices, y, train_mask, test_mask, G_undir, last_nodes, target_nodes = sdg.load_dataset(folder)
        B1, B2 = B_matrices
        target_nodes_all.append(target_nodes)
        inputs_all.append([None, onp.array(last_nodes), X])
        y_all.append(y)
    model = HYPERPARAMS['model']
    if model not in ('scone', 'ebli', 'bunch'):
        raise Exception('invalid model type')
    cx = SimplicialComplex.from_dense(B1, B2, model if model != 'bunch' else 'scone', flips=flips)
    if model == 'bunch':
        shifts = [CsrOperator(M) for M in compute_shift_matrices(B1, B2)]      # trajectory_experiments.py:255-257
    else:
        shifts = shift_handles(cx)

    e = onp.nonzero(B1.T)[1]
    edges = onp.array_split(e, len(e) / 2)
    E, E_lookup = [], {}
    for i, e in enumerate(edges):
        E.append(tuple(e))
        E_lookup[tuple(e)] = i

    last_nodes = inputs_all[0][1]
    n_nbrs = onp.array([len(G_undir[n]) for n in last_nodes])
    nbrhoods = onp.array(cx.nbrhoods)
    try:
        prefixes = list(onp.load('trajectory_data_1hop_' + folder_suffix + '/prefixes.npy', allow_pickle=True))
    except Exception:
        prefixes = [sdg.flow_to_path(inputs_all[0][-1][i], E, last_nodes[i]) for i in range(len(last_nodes))]

    if flips is not None:
        for i in range(len(inputs_all)):
            n_flows, n_edges = inputs_all[i][-1].shape[:2]
            inputs_all[i][-1] = (inputs_all[i][-1].reshape((n_flows, n_edges)) * flips[None, :]).reshape((n_flows, n_edges, 1))
    Bconds_func = Bconds(cx)
    for i in range(len(inputs_all)):
        inputs_all[i][0] = Bconds_func if model != 'bunch' else nbrhoods
    return inputs_all, y_all, train_mask, test_mask, shifts, G_undir, E_lookup, nbrhoods, n_nbrs, target_nodes_all, prefixes


def train_model():
    """
    Trains a model to predict the next node in each input path           (trajectory_experiments.py:313-513)
    """
    # launched under torchrun (WORLD_SIZE > 1): one process per GPU, the batch of every step is sharded over the ranks and the
    # weight gradients all-reduced (Scone_GCN.train); a plain `python trajectory_experiments.py ...` run is unchanged
    from . import dp as _dp
    _dp.init_from_env()
    inputs_all, y_all, train_mask, test_mask, shifts, G_undir, E_lookup, nbrhoods, n_nbrs, target_nodes_all, prefixes = \
        data_setup(hops=(1, 2), load=HYPERPARAMS['load_data'], folder_suffix=HYPERPARAMS['data_folder_suffix'])
    (inputs_1hop, inputs_2hop), (y_1hop, y_2hop) = inputs_all, y_all
    in_axes = tuple(([None] * len(shifts)) + [None, None, 0, 0])
    if HYPERPARAMS['markov'] == 1:
        raise NotImplementedError('the Markov baseline is outside the accelerated path (SURVEY.md §2)')

    scone = Scone_GCN(HYPERPARAMS['epochs'], HYPERPARAMS['learning_rate'], HYPERPARAMS['batch_size'],
                      HYPERPARAMS['weight_decay'])
    model_func = {'scone': scone_func, 'ebli': ebli_func, 'bunch': bunch_func}.get(HYPERPARAMS['model'])
    if model_func is None:
        raise Exception('invalid model')
    scone.setup(model_func, HYPERPARAMS['hidden_layers'], shifts, inputs_1hop, y_1hop, in_axes, train_mask,
                model_type=HYPERPARAMS['model'])

    if HYPERPARAMS['regional']:
        train_mask = onp.array([1 if i % 3 == 1 else 0 for i in range(len(y_1hop))])
        test_mask = onp.array([1 if i % 3 == 2 else 0 for i in range(len(y_1hop))])

    if HYPERPARAMS['describe'] == 1:
        print('Graph nodes: {}, edges: {}, avg degree: {}'.format(
            len(G_undir.nodes), len(G_undir.edges), onp.average([G_undir.degree[node] for node in G_undir.nodes])))
        print('Training paths: {}, Test paths: {}'.format(train_mask.sum(), test_mask.sum()))
        print('Model: {}'.format(HYPERPARAMS['model']))

    if HYPERPARAMS['load_model']:
        scone.weights = list(onp.load('models/' + HYPERPARAMS['model_name'] + '.npy', allow_pickle=True))
        if HYPERPARAMS['epochs'] != 0:
            scone.train(inputs_1hop, y_1hop, train_mask, test_mask, n_nbrs)
            os.makedirs('models', exist_ok=True)
            onp.save('models/' + HYPERPARAMS['model_name'], onp.array(scone.weights, dtype=object), allow_pickle=True)
        (train_loss, train_acc), (test_loss, test_acc) = scone.test(inputs_1hop, y_1hop, train_mask, n_nbrs), \
            scone.test(inputs_1hop, y_1hop, test_mask, n_nbrs)
    else:
        train_loss, train_acc, test_loss, test_acc = scone.train(inputs_1hop, y_1hop, train_mask, test_mask, n_nbrs)
        os.makedirs('models', exist_ok=True)
        onp.save('models/' + HYPERPARAMS['model_name'], onp.array(scone.weights, dtype=object), allow_pickle=True)

    print('standard test set:')
    train_2target, test_2target = scone.two_target_accuracy(shifts, inputs_1hop, y_1hop, train_mask, n_nbrs), \
        scone.two_target_accuracy(shifts, inputs_1hop, y_1hop, test_mask, n_nbrs)
    scone.test(inputs_1hop, y_1hop, test_mask, n_nbrs)
    print('2-target accs:', train_2target, test_2target)

    if HYPERPARAMS['reverse']:
        sfx = HYPERPARAMS['data_folder_suffix']
        rev_flows_in, rev_targets_1hop, rev_last_nodes = \
            onp.load('trajectory_data_1hop_' + sfx + '/rev_flows_in.npy'), \
            onp.load('trajectory_data_1hop_' + sfx + '/rev_targets.npy'), \
            onp.load('trajectory_data_1hop_' + sfx + '/rev_last_nodes.npy')
        rev_n_nbrs = [len(sdg.neighborhood(G_undir, n)) for n in rev_last_nodes]
        print('Reverse experiment:')
        scone.test([inputs_1hop[0], rev_last_nodes, rev_flows_in], rev_targets_1hop, test_mask, rev_n_nbrs)
    return scone, (train_loss, train_acc, test_loss, test_acc)


if __name__ == '__main__':
    train_model()
