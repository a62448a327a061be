"""Minimal NumPy-backed stand-in for the `jax` API surface that jax_cosmo uses.

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference source
under /root/reference can be executed in float64 in a container without JAX,
to generate golden vectors (oracle/make_golden.py).  It is deliberately dumb:
`vmap`, `lax.map` and `lax.scan` are Python loops, `jit` is the identity.
Nothing in the product path imports this.
"""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, tree_util  # noqa: F401


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def jit(fn=None, static_argnums=None, static_argnames=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def _tree_stack(outs, axis):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(
            _tree_stack([o[i] for o in outs], axis) for i in range(len(first))
        )
    return _np.stack([_np.asarray(o) for o in outs], axis=axis)


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            call = [
                a if ax is None else _np.take(_np.asarray(a), i, axis=ax)
                for a, ax in zip(args, axes)
            ]
            outs.append(fn(*call))
        return _tree_stack(outs, out_axes)

    return mapped


def stop_gradient(x):
    return x
