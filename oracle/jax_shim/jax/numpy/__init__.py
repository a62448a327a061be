"""jax.numpy stand-in: plain NumPy float64, plus jnp.clip's one-bound call form."""
import numpy as _np

globals().update({k: getattr(_np, k) for k in dir(_np) if not k.startswith("__")})
from numpy import linalg  # noqa: F401,E402


def clip(x, min=None, max=None, **kw):  # jnp.clip(x, lo) is legal in JAX
    if "a_min" in kw:
        min = kw["a_min"]
    if "a_max" in kw:
        max = kw["a_max"]
    return _np.clip(x, min, max)


