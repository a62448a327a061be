def odeint(*a, **k):  # imported by jax_cosmo/core.py, never called on the hot path
    raise NotImplementedError
