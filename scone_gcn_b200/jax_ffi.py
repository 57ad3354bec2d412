"""jax.ffi binding of the model-level C ABI ("XLA FFI (outer)") — import-guarded.

The reference is JAX host code: `Scone_GCN.setup` vmaps the per-sample model (scone_trajectory_model.py:256), `loss` /
`accuracy` call it (:46,:64) and `grad(self.loss)` differentiates w.r.t. the weight list (:307).  With jax installed this
module lets that code keep its arrays in JAX:

    fns = jax_ffi.bind(net)                     # net: scone_gcn_b200.SconeModel
    lp  = fns.logprobs(w_flat, ptr, edge, val, last)                       # [B, D]      == vmap(scone_func)(...)
    nll = fns.nll_sum(w_flat, ptr, edge, val, last, target_idx, mask)      # scalar, differentiable w.r.t. w_flat
    loss = nll / mask.sum() + wd * (w_flat ** 2).sum()                     # scone_trajectory_model.py:42-56
    g = jax.grad(...)(w_flat)                                              # one fused forward + backward on the GPU

STATUS: jax / jaxlib are NOT installable in this repository's build image (no wheel, no network), so nothing below the
import guard has been executed there; `csrc/scone_xla_ffi.cc` holds the handlers and is compiled by
`__graft_entry__.build()` only where jaxlib's `xla/ffi/api/ffi.h` exists.  API names follow jax >= 0.5 (`jax.ffi`);
0.4.3x has the same functions under `jax.extend.ffi`.
"""
import ctypes
import os

try:                                     # import guard: the package itself never needs jax
    import jax
    import jax.numpy as jnp
    _ffi = getattr(jax, 'ffi', None)
    if _ffi is None:
        from jax.extend import ffi as _ffi
    _IMPORT_ERROR = None
except Exception as exc:                 # pragma: no cover - jax is absent in the build image
    jax = jnp = _ffi = None
    _IMPORT_ERROR = exc

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libscone_b200_xla.so')
_TARGETS = {'scone_model_forward': 'SconeModelForward', 'scone_model_loss_grad': 'SconeModelLossGrad',
            'scone_accuracy': 'SconeAccuracy'}
_registered = False


class JaxUnavailable(RuntimeError):
    pass


def available():
    """True when jax imports and the FFI shim library was built (jaxlib headers were present at build time)."""
    return jax is not None and os.path.exists(_SHIM)


def register():
    """jax.ffi.register_ffi_target for every handler of csrc/scone_xla_ffi.cc (platform CUDA).  Idempotent."""
    global _registered
    if jax is None:
        raise JaxUnavailable('jax is not importable here (%r); use the Python host mirror (scone_gcn_b200.Scone_GCN) or the C ABI' % (_IMPORT_ERROR,))
    if not os.path.exists(_SHIM):
        raise JaxUnavailable('%s was not built: __graft_entry__.build() compiles csrc/scone_xla_ffi.cc only where jaxlib ships '
                             'xla/ffi/api/ffi.h' % _SHIM)
    if _registered:
        return
    from . import _lib
    _lib.lib()                                              # libscone_b200.so first: the shim links against it
    shim = ctypes.CDLL(_SHIM, mode=ctypes.RTLD_GLOBAL)
    if not shim.scone_xla_ffi_available():
        raise JaxUnavailable('the FFI shim was compiled without xla/ffi/api/ffi.h')
    for name, sym in _TARGETS.items():
        _ffi.register_ffi_target(name, _ffi.pycapsule(getattr(shim, sym)), platform='CUDA')
    _registered = True


class BoundModel:
    """JAX-callable views of one SconeModel (the handle is passed as an int64 attribute; the model must outlive the calls)."""

    def __init__(self, net):
        register()
        self.net = net
        self.handle = int(ctypes.cast(net.handle, ctypes.c_void_p).value)
        self.D = int(net.cx.D)
        self.n_params = int(net.n_params)

        @jax.custom_vjp
        def nll_sum(w_flat, ptr, edge, val, last, target_idx, mask):
            return self._loss_grad(w_flat, ptr, edge, val, last, target_idx, mask)[self.n_params]

        def fwd(w_flat, ptr, edge, val, last, target_idx, mask):
            buf = self._loss_grad(w_flat, ptr, edge, val, last, target_idx, mask)
            return buf[self.n_params], buf[:self.n_params]             # residual: d nll_sum / d w_flat, from the same pass

        def bwd(grads, ct):
            return (ct * grads, None, None, None, None, None, None)    # the reference differentiates w.r.t. the weights only (:307)
        nll_sum.defvjp(fwd, bwd)
        self.nll_sum = nll_sum

    def logprobs(self, w_flat, ptr, edge, val, last):
        """[B, D] log-probs == vmap(scone_func / ebli_func)(weights, *shifts, Bconds, last, flows)[:, :, 0]."""
        out = jax.ShapeDtypeStruct((last.shape[0], self.D), jnp.float32)
        return _ffi.ffi_call('scone_model_forward', out)(w_flat.astype(jnp.float32), ptr, edge, val, last, model=self.handle)

    def _loss_grad(self, w_flat, ptr, edge, val, last, target_idx, mask):
        out = jax.ShapeDtypeStruct((self.n_params + 2,), jnp.float32)  # [grads | nll_sum | count], unnormalised sums
        return _ffi.ffi_call('scone_model_loss_grad', out)(w_flat.astype(jnp.float32), ptr, edge, val, last, target_idx,
                                                           mask.astype(jnp.float32), model=self.handle)

    def accuracy(self, logprobs, n_nbrs, target_idx, mask):
        """(correct, counted) as an int32[2] array — scone_trajectory_model.py:59-71 on the device."""
        out = jax.ShapeDtypeStruct((2,), jnp.int32)
        return _ffi.ffi_call('scone_accuracy', out)(logprobs, n_nbrs, target_idx, mask.astype(jnp.float32))

    def flatten(self, weights):
        """The reference's weight LIST (generate_weights order, :215-242) -> the flat vector the ops take."""
        return jnp.concatenate([jnp.ravel(jnp.asarray(w, jnp.float32)) for w in weights])

    def loss(self, weights, ptr, edge, val, last, target_idx, mask, weight_decay):
        """Scone_GCN.loss (:42-56): masked mean NLL + ridge; jax.grad of this w.r.t. `weights` runs one fused GPU pass."""
        w_flat = self.flatten(weights)
        return self.nll_sum(w_flat, ptr, edge, val, last, target_idx, mask) / jnp.sum(mask) + weight_decay * jnp.sum(w_flat ** 2)


def bind(net):
    return BoundModel(net)
