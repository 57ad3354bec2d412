"""Kernel-level timing probe (CUDA events) for the fused layer kernels on a synthetic complex.
Usage: python tools/probe_perf.py [n_nodes] [b] [C]"""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scone_gcn_b200 as sg
from scone_gcn_b200 import _lib
from scone_gcn_b200 import synthetic_data_gen as sdg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
C = int(sys.argv[3]) if len(sys.argv) > 3 else 32
t0 = time.time()
coords, valid, faces, edges = sdg._complex_arrays(n)
print('complex: N=%d E=%d F=%d  (%.1fs)' % (n, len(edges), len(faces), time.time() - t0), flush=True)
t0 = time.time()
cx = sg.SimplicialComplex.from_simplices(n, edges, faces, 'scone')
print('handle built in %.1fs: D=%d nnz=%s' % (time.time() - t0, cx.D, cx.nnz), flush=True)
L = _lib.lib()
E = cx.E
dev = torch.device('cuda')
st = torch.cuda.current_stream().cuda_stream
PEAK = 6554.2


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print('%-34s %9.3f ms  %8.1f GB/s algorithmic  (%.1f%% of %.0f)' % (name, ms, gbs, 100 * gbs / PEAK, PEAK), flush=True)


W = [torch.randn(C, C, device=dev) * 0.2 for _ in range(3)]
W1 = [torch.randn(1, C, device=dev) * 0.2 for _ in range(3)]
for kind in ('dense-random', 'rows-2pct+flags', 'blob-1pct+flags'):
    H = torch.randn(E, b, C, device=dev)
    if kind.startswith('rows'):
        H *= (torch.rand(E, 1, 1, device=dev) < 0.02)
    if kind.startswith('blob'):          # spatially compact support, different per trajectory (what real activations look like)
        mid = torch.empty(E, 2)
        mid[torch.from_numpy(cx.edge_rank.astype(np.int64))] = torch.from_numpy(coords[edges].mean(axis=1)).float()
        mid = mid.to(dev)
        ctr = torch.rand(b, 2, device=dev)
        H *= ((mid[:, None, :] - ctr[None, :, :]).norm(dim=2) < 0.056)[:, :, None]
    occ = (H.abs().amax(dim=2) > 0).to(torch.uint8).contiguous() if 'flags' in kind else None
    X = H[:, :, 0].contiguous() if 'flags' in kind else torch.randn(E, b, device=dev)
    occ_o = torch.empty(E, b, dtype=torch.uint8, device=dev)
    scr = torch.empty(L.scone_occ_scratch_bytes(cx.handle, b), dtype=torch.uint8, device=dev) if occ is not None else None
    occX = (X != 0).to(torch.uint8).contiguous() if occ is not None else None
    out = torch.empty_like(H)
    ms = timeit(lambda: _lib.check(L.scone_layer_forward(cx.handle, 0, b, C, C, _lib.dptr(H), _lib.dptr(W[0]), _lib.dptr(W[1]),
                                                        _lib.dptr(W[2]), _lib.dptr(out), _lib.dptr(occ), _lib.dptr(occ_o), _lib.dptr(scr), st)))
    report('fwd %d->%d (%s)' % (C, C, kind), ms, 4 * E * b * 2 * C)
    ws = torch.empty(L.scone_layer_backward_workspace_bytes(C, C) // 4 + 16, device=dev)
    dW = torch.zeros(3, C, C, device=dev)
    ms = timeit(lambda: _lib.check(L.scone_layer_backward(cx.handle, 0, b, C, C, _lib.dptr(H), _lib.dptr(out), _lib.dptr(W[0]),
                                                         _lib.dptr(W[1]), _lib.dptr(W[2]), _lib.dptr(out), _lib.dptr(dW), 0,
                                                         _lib.dptr(ws), _lib.dptr(occ), _lib.dptr(occ), _lib.dptr(occ_o), _lib.dptr(scr), st)))
    report('bwd %d->%d (%s)' % (C, C, kind), ms, 4 * E * b * 3 * C)
    ms = timeit(lambda: _lib.check(L.scone_layer_forward(cx.handle, 0, b, 1, C, _lib.dptr(X), _lib.dptr(W1[0]), _lib.dptr(W1[1]),
                                                        _lib.dptr(W1[2]), _lib.dptr(out), _lib.dptr(occX), _lib.dptr(occ_o), _lib.dptr(scr), st)))
    report('fwd 1->%d (%s)' % (C, kind), ms, 4 * E * b * (1 + C))
    ms = timeit(lambda: _lib.check(L.scone_layer_backward(cx.handle, 0, b, 1, C, _lib.dptr(H), _lib.dptr(X), None, None, None,
                                                         None, _lib.dptr(dW), 0, _lib.dptr(ws), _lib.dptr(occ), None, None, _lib.dptr(scr), st)))
    report('bwd 1->%d (%s)' % (C, kind), ms, 4 * E * b * (1 + C))
    del H, out
a = torch.empty(E * b * C, device=dev)
c = torch.empty_like(a)
ms = timeit(lambda: c.copy_(a))
report('torch copy (reference point)', ms, 8 * a.numel())
