"""Redshift <-> scale factor conversions with the reference's names (jax_cosmo/utils.py:2-9); work on floats and arrays alike."""

__all__ = ["z2a", "a2z"]


def z2a(z):
    """Scale factor a = 1 / (1 + z)."""
    one_plus_z = 1.0 + z
    return 1.0 / one_plus_z


def a2z(a):
    """Redshift z = 1 / a - 1."""
    inv_a = 1.0 / a
    return inv_a - 1.0
