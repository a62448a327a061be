"""jax.lax stand-in (Python loops)."""
import numpy as _np


def map(f, xs):  # noqa: A001
    return _np.stack([_np.asarray(f(x)) for x in xs], axis=0)


def scan(f, init, xs):
    carry = init
    ys = []
    for x in xs:
        carry, y = f(carry, x)
        ys.append(y)
    first = ys[0]
    if isinstance(first, (tuple, list)):
        ys = type(first)(_np.stack([y[i] for y in ys]) for i in range(len(first)))
    else:
        ys = _np.stack([_np.asarray(y) for y in ys])
    return carry, ys


def switch(index, branches, *operands):
    return branches[int(index)](*operands)
