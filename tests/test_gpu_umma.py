"""The tcgen05 / TMEM dense layer kernel (csrc/scone_umma.cu, scone_set_dense_kernel(3)) through the C ABI against the fp32 SIMT dense
kernel on the same input: same gather order, 3xTF32 product within fp32 rounding noise; complexes whose edge count is not a
multiple of the 8-edge tile, several trajectory-slab counts, all three activations."""
import numpy as np
import pytest
import torch

from golden_util import Dataset

pytestmark = pytest.mark.gpu


def _run(cx, act, b, H, W, which):
    from scone_gcn_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    L.scone_set_dense_kernel(which)
    try:
        out = torch.full((cx.E, b, 32), 7.0, device=H.device)
        _lib.check(L.scone_layer_forward(cx.handle, act, b, 32, 32, _lib.dptr(H), _lib.dptr(W[0]), _lib.dptr(W[1]), _lib.dptr(W[2]),
                                         _lib.dptr(out), None, None, None, st))
        if which == 3:
            _lib.check(L.scone_umma_status(st), 'scone_umma_status')
        return out.cpu().numpy()
    finally:
        L.scone_set_dense_kernel(1)


@pytest.mark.parametrize('act', [0, 1, 2])
@pytest.mark.parametrize('b', [16, 48, 64])
def test_umma_forward_matches_simt_dense_kernel(act, b):
    import scone_gcn_b200 as sg
    ds = Dataset('dataset_small.npz')
    cx = sg.SimplicialComplex.from_dense(ds.B1, ds.B2, 'scone')
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(100 + 7 * act + b)
    H = torch.randn(cx.E, b, 32, generator=g).to(dev)
    W = [(torch.randn(32, 32, generator=g) * 0.3).to(dev) for _ in range(3)]
    ref = _run(cx, act, b, H, W, 0)
    got = _run(cx, act, b, H, W, 3)
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()


def test_umma_forward_on_a_larger_sparse_complex_many_tiles_per_cta():
    """~9000 edges x 64 trajectories: hundreds of tiles, both warp groups of every CTA cycle through many mbarrier phases."""
    import scone_gcn_b200 as sg
    from scone_gcn_b200 import synthetic_data_gen as sdg
    sp = sdg.generate_sparse_dataset(3000, 8, seed=5, n_waypoints=8)
    cx = sg.SimplicialComplex.from_simplices(int(sp.n_nodes), sp.edges, sp.faces, 'scone')
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(3)
    b = 64
    H = torch.randn(cx.E, b, 32, generator=g).to(dev)
    W = [(torch.randn(32, 32, generator=g) * 0.2).to(dev) for _ in range(3)]
    ref = _run(cx, 0, b, H, W, 1)                          # slab kernel (mma.sync 3xTF32): same gather, same split
    got = _run(cx, 0, b, H, W, 3)
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    got2 = _run(cx, 0, b, H, W, 3)
    assert np.array_equal(got, got2)                       # run-to-run bit-exact
