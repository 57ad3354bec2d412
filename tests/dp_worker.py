"""Worker of tests/test_gpu_dp.py: launched by torchrun with one rank per GPU; trains the product Scone_GCN data-parallel and
has rank 0 save the all-reduced gradient buffer of one batch and the weights after a few Adam steps."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def run(out, data_parallel):
    import torch
    import scone_gcn_b200 as sg
    from scone_gcn_b200 import dp
    from scone_gcn_b200.scone_trajectory_model import Scone_GCN
    from scone_gcn_b200 import trajectory_experiments as te
    from golden_util import Dataset, load, weights_of
    rank, world = dp.init_from_env() if data_parallel else (0, 1)
    ds = Dataset('dataset_small.npz')
    fx = load('model_small_scone_h16.npz')
    cx = sg.SimplicialComplex.from_dense(ds.B1, ds.B2, 'scone')
    inputs = [te.Bconds(cx), ds.last_nodes, ds.flows]
    np.random.seed(1030)
    net = Scone_GCN(3, 1e-3, 16, 5e-5, verbose=False, data_parallel=data_parallel)
    net.setup(te.scone_func, [(3, 16)] * 3, te.shift_handles(cx), inputs, ds.targets, None, ds.train_mask)
    net.weights = weights_of(fx, 'w_big')
    # one gradient over the whole training set, sharded by hand exactly as train() shards a batch
    rows = np.nonzero(ds.train_mask == 1)[0]
    mine = dp.shard_rows(rows) if data_parallel else rows
    p = net._prepared(inputs)
    net._push(net.weights)
    ptr, fe, fv, last = p.select(mine)
    net._net.loss_grad(ptr, fe, fv, last, ds.raw['targets_argmax'][mine].astype(np.int32), np.ones(len(mine), np.float32), read=False)
    ident = True
    if data_parallel:
        if os.environ.get('SCONE_DP_EXCHANGE', 'peer') == 'peer':
            # the fused exchange: sum over ranks (left in the gradient buffer) + one Adam step on a scratch copy of the weights
            ex = dp.make_exchange(net._net.n_params + 2, torch.device('cuda', torch.cuda.current_device()))
            assert ex is not None, 'peer exchange could not be set up'
            w_before = net._net.get_weights()
            ex.adam_step(net._net, 0, 1e-3, 5e-5)
            ex.status()
            wt = net._net.weights_tensor().clone()
            both = [torch.empty_like(wt) for _ in range(world)]
            torch.distributed.all_gather(both, wt)
            ident = all(bool(torch.equal(both[0], b)) for b in both[1:])       # rank-ordered sums: bit-identical on every rank
            net._net.set_weights(w_before, reset_adam=True)
            ex.close()
        else:
            dp.allreduce_sum_(net._net.grads_tensor())
    grads = net._net.read_grads()
    n_nbrs = fx['n_nbrs']
    res = net.train(inputs, ds.targets, ds.train_mask, ds.test_mask, n_nbrs)
    if rank == 0:
        np.savez(out, grads=grads, result=np.asarray(res, np.float64), world=world, ident=ident, **{'w%d' % i: w for i, w in enumerate(net.weights)})
    if data_parallel:
        torch.cuda.synchronize()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    run(sys.argv[1], data_parallel=os.environ.get('WORLD_SIZE', '1') != '1')
