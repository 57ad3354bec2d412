"""jax.experimental.optimizers.adam stand-in (reference call site: scone_trajectory_model.py:11,300).

Formula restated from upstream JAX (jax/example_libraries/optimizers.py, `adam`), which is NOT under
/root/reference: m = (1-b1) g + b1 m ; v = (1-b2) g^2 + b2 v ;
mhat = m/(1-b1^(i+1)) ; vhat = v/(1-b2^(i+1)) ; x = x - step_size * mhat / (sqrt(vhat) + eps),
b1=0.9, b2=0.999, eps=1e-8.  Powers b^(i+1) are evaluated in the array dtype (fp32), as upstream's
`jnp.asarray(b1, m.dtype) ** (i + 1)` does.
"""
import torch as _torch
from ..numpy import _as_tensor


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
    def init(x0):
        xs = [_as_tensor(x).detach().clone() for x in x0]
        return [(x, _torch.zeros_like(x), _torch.zeros_like(x)) for x in xs]

    def update(i, g, state):
        new = []
        for gi, (x, m, v) in zip(g, state):
            gi = gi.detach().as_subclass(_torch.Tensor)
            m = (1 - b1) * gi + b1 * m
            v = (1 - b2) * gi * gi + b2 * v
            one = _torch.ones((), dtype=x.dtype)
            mhat = m / (one - _torch.tensor(b1, dtype=x.dtype) ** (i + 1))
            vhat = v / (one - _torch.tensor(b2, dtype=x.dtype) ** (i + 1))
            x = x - step_size * mhat / (_torch.sqrt(vhat) + eps)
            new.append((x, m, v))
        return new

    def get_params(state):
        return [x for (x, m, v) in state]

    return init, update, get_params
