"""Galaxy / IA bias descriptors with the reference's constructors (jax_cosmo/bias.py:10-57);
evaluated on the 513 Limber nodes by the CUDA path."""
from jax_cosmo_b200.jax_utils import container

__all__ = ["constant_linear_bias", "inverse_growth_linear_bias", "des_y1_ia_bias"]


class constant_linear_bias(container):
    """b(z) = b (bias.py:10-22)."""
    _family = "constant"


class inverse_growth_linear_bias(container):
    """b(z) = b / D(a) (bias.py:25-39)."""
    _family = "inverse_growth"


class des_y1_ia_bias(container):
    """b(z) = A ((1+z)/(1+z0))^eta (bias.py:42-57)."""
    _family = "des_y1_ia"
