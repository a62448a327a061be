"""jax_cosmo/utils.py:2-9."""


def z2a(z):
    return 1.0 / (1.0 + z)


def a2z(a):
    return 1.0 / a - 1.0
