"""Forward tcgen05 dense kernel: time against the tile-chunk size (scone_set_dense_chunk).  Usage: python tools/sweep_dense_chunk.py [n_nodes]"""
import json, os, sys
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
L.scone_set_dense_kernel(3)
ref = None
for chunk in (0, 1, 2, 4, 8, 16, 32, 64):
    L.scone_set_dense_chunk(chunk)
    def fn():
        _lib.check(L.scone_layer_forward(cx.handle, 0, b, C, C, _lib.dptr(H), _lib.dptr(W[0]), _lib.dptr(W[1]), _lib.dptr(W[2]), _lib.dptr(O), None, None, None, st))
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(4):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(z))
    _lib.check(L.scone_umma_status(st), 'umma status')
    if ref is None:
        ref = O.clone()
    same = bool(torch.equal(ref, O))
    print('chunk %3d: %.3f ms  %.3f of 6554 GB/s   identical to chunk 0: %s' % (chunk, best, 4.0 * E * b * 2 * C / best / 1e6 / 6554.2, same), flush=True)
L.scone_set_dense_chunk(0)
L.scone_set_dense_kernel(_lib.DEFAULT_DENSE_KERNEL)
