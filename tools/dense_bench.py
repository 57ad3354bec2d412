"""Dense-tile roofline of one fused 32 -> 32 layer forward on [E][64][32] random features: the dense kernels side by side
(1 = slab / mma.sync, 3 = tcgen05 tiles).  Usage: python tools/dense_bench.py [n_nodes]"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scone_gcn_b200 as sg
from scone_gcn_b200 import _lib, synthetic_data_gen as sdg

n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 370000
sp = sdg.generate_sparse_dataset(n_nodes, 8, seed=1030, n_waypoints=8)
cx = sg.SimplicialComplex.from_simplices(int(sp.n_nodes), sp.edges, sp.faces, 'scone')
L = _lib.lib()
dev = torch.device('cuda')
E, b, C = cx.E, 64, 32
H = torch.randn(E, b, C, device=dev)
O = torch.empty_like(H)
W = [torch.randn(C, C, device=dev) * 0.2 for _ in range(3)]
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
res = {}
outs = {}
for which in (1, 3):
    L.scone_set_dense_kernel(which)
    def fn():
        _lib.check(L.scone_layer_forward(cx.handle, 0, b, C, C, _lib.dptr(H), _lib.dptr(W[0]), _lib.dptr(W[1]), _lib.dptr(W[2]), _lib.dptr(O), None, None, None, st))
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(z))
    if which == 3:
        _lib.check(L.scone_umma_status(st), 'umma status')
    outs[which] = O[:4096].clone()
    gbs = 4.0 * E * b * 2 * C / best / 1e6
    res[which] = dict(ms=best, gbs=gbs, frac=gbs / peak)
L.scone_set_dense_kernel(_lib.DEFAULT_DENSE_KERNEL)
print(json.dumps({'E': E, 'b': b, 'C': C, 'peak_gbs': peak, 'slab_mma_sync': res[1], 'umma_tcgen05': res[3],
                  'max_abs_diff_first_rows': float((outs[1] - outs[3]).abs().max())}))
