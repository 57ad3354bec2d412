"""Dense fused-layer timing probe: fp32 SIMT tile kernels (0) vs slab kernels (1: 16x1, 2: 8x2) on dense random features.
Usage: python tools/probe_dense.py [n_nodes] [b] [C]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scone_gcn_b200 as sg
from scone_gcn_b200 import _lib
from scone_gcn_b200 import synthetic_data_gen as sdg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Cs = [int(sys.argv[3])] if len(sys.argv) > 3 else [32, 16]
coords, valid, faces, edges = sdg._complex_arrays(n)
cx = sg.SimplicialComplex.from_simplices(n, edges, faces, 'scone')
L = _lib.lib()
E = cx.E
dev = torch.device('cuda')
st = torch.cuda.current_stream().cuda_stream
try:
    import json
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    PEAK = 6554.2
print('complex: N=%d E=%d F=%d D=%d nnz=%s b=%d' % (n, E, len(faces), cx.D, cx.nnz, b), flush=True)


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


for C in Cs:
    H = torch.randn(E, b, C, device=dev)
    out = torch.empty_like(H)
    G = torch.randn(E, b, C, device=dev)
    Gp = torch.empty_like(H)
    ws = torch.empty(L.scone_layer_backward_workspace_bytes(C, C) // 4 + 16, device=dev)
    dW = torch.zeros(3, C, C, device=dev)
    for wscale in (0.2, 0.01):
        W = [torch.randn(C, C, device=dev) * wscale for _ in range(3)]
        ref = None
        for which in (0, 1, 2):
            L.scone_set_dense_kernel(which)
            ms = timeit(lambda: _lib.check(L.scone_layer_forward(cx.handle, 0, b, C, C, _lib.dptr(H), _lib.dptr(W[0]), _lib.dptr(W[1]),
                                                                _lib.dptr(W[2]), _lib.dptr(out), None, None, None, st)))
            gbs = 4.0 * E * b * 2 * C / ms / 1e6
            if ref is None:
                ref = out.clone()
            err = float((out - ref).abs().max())
            print('fwd C=%d wscale=%.2f kernel=%d: %8.3f ms  %7.1f GB/s algorithmic  %.1f%% of %.0f   max|diff vs kernel 0| %.2e'
                  % (C, wscale, which, ms, gbs, 100 * gbs / PEAK, PEAK, err), flush=True)
            msb = timeit(lambda: _lib.check(L.scone_layer_backward(cx.handle, 0, b, C, C, _lib.dptr(G), _lib.dptr(H), _lib.dptr(W[0]),
                                                                 _lib.dptr(W[1]), _lib.dptr(W[2]), _lib.dptr(Gp), _lib.dptr(dW), 0,
                                                                 _lib.dptr(ws), None, None, None, None, st)))
            gbs = 4.0 * E * b * 3 * C / msb / 1e6
            print('bwd C=%d wscale=%.2f kernel=%d: %8.3f ms  %7.1f GB/s algorithmic  %.1f%% of %.0f' % (C, wscale, which, msb, gbs, 100 * gbs / PEAK, PEAK),
                  flush=True)
    L.scone_set_dense_kernel(_lib.DEFAULT_DENSE_KERNEL)
    del H, out, G, Gp
a = torch.empty(E * b * 32, device=dev)
c = torch.empty_like(a)
ms = timeit(lambda: c.copy_(a))
print('torch copy reference: %.3f ms %.1f GB/s' % (ms, 8 * a.numel() / ms / 1e6))
