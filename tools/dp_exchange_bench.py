"""Latency of the per-step gradient exchange at N ranks: the library's NVLink peer-memory kernel fused with Adam (scone_dp.cu)
against NCCL all-reduce + Adam kernel, on the cfg5 payload (6,276 floats), (a) back to back and (b) inside a small training step.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_exchange_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist

import scone_gcn_b200 as sg
from scone_gcn_b200 import _lib, dp
from golden_util import Dataset

rank, world = dp.init_from_env()
dev = torch.device('cuda', torch.cuda.current_device())
L = _lib.lib()
ds = Dataset('dataset_default.npz')
cx = sg.SimplicialComplex.from_dense(ds.B1, ds.B2, 'scone')
net = sg.SconeModel(cx, [32, 32, 32], micro_batch=1024)
rs = np.random.RandomState(1030)
net.set_weights([0.01 * rs.randn(*s) for s in net.shapes])
from scone_gcn_b200.complex import flows_to_csr
ptr, fe, fv = flows_to_csr(np.asarray(ds.flows))
n = len(ptr) - 1
lo, hi = dp.shard_range(n, rank, world)
p0, p1 = int(ptr[lo]), int(ptr[hi])
d = dict(ptr=torch.from_numpy((ptr[lo:hi + 1] - ptr[lo]).astype(np.int32)).to(dev), edge=torch.from_numpy(fe[p0:p1].astype(np.int32)).to(dev),
         val=torch.from_numpy(fv[p0:p1].astype(np.float32)).to(dev), last=torch.from_numpy(np.asarray(ds.last_nodes)[lo:hi].astype(np.int32)).to(dev),
         tgt=torch.from_numpy(ds.raw['targets_argmax'][lo:hi].astype(np.int32)).to(dev), mask=torch.ones(hi - lo, device=dev))
B = hi - lo
stream = torch.cuda.current_stream().cuda_stream
gbuf = net.grads_tensor()
ex = dp.make_exchange(net.n_params + 2, dev)
step = [0]


def grad():
    _lib.check(L.scone_model_loss_grad_dev(net.handle, B, _lib.dptr(d['ptr']), _lib.dptr(d['edge']), _lib.dptr(d['val']), _lib.dptr(d['last']),
                                           _lib.dptr(d['tgt']), _lib.dptr(d['mask']), 1, stream))


def opt(mode):
    if mode == 'peer':
        ex.adam_step(net, step[0], 1e-3, 5e-5, stream)
    else:
        dist.all_reduce(gbuf)
        net.adam_step(step[0], 1e-3, 5e-5, stream)
    step[0] += 1


def timed(fn, iters):
    for _ in range(10):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    z.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(z) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


grad()
out = {'world': world, 'payload_floats': net.n_params + 2, 'peer_available': ex is not None}
modes = (['peer'] if ex is not None else []) + ['nccl']
for mode in modes:
    out['exchange_only_us_' + mode] = 1e3 * timed(lambda: opt(mode), 300)
    out['small_step_us_' + mode] = 1e3 * timed(lambda: (grad(), opt(mode)), 200)
if ex is not None:
    ex.status(stream)
out['grad_only_us'] = 1e3 * timed(grad, 200)
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize()
dist.destroy_process_group()
