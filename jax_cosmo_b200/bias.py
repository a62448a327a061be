"""Galaxy / IA bias descriptors with the reference's constructors (jax_cosmo/bias.py:10-57).  Inside `angular_cl` they
are evaluated on the 513 Limber nodes by the CUDA path; called directly, `bias(cosmo, z)`, they return the bias at z like
the reference: the two redshift-only forms are one-line host expressions, the growth-dependent one takes D(a) from the
device (`background.growth_factor`, no CPU fallback)."""
import numpy as np

from jax_cosmo_b200.jax_utils import container

__all__ = ["constant_linear_bias", "inverse_growth_linear_bias", "des_y1_ia_bias"]


class constant_linear_bias(container):
    """b(z) = b (bias.py:10-22)."""
    _family = "constant"

    def __call__(self, cosmo, z):
        return self.params[0] * np.ones_like(np.asarray(z, dtype=np.float64))


class inverse_growth_linear_bias(container):
    """b(z) = b / D(a) (bias.py:25-39)."""
    _family = "inverse_growth"

    def __call__(self, cosmo, z):
        from jax_cosmo_b200 import background
        return self.params[0] / background.growth_factor(cosmo, 1.0 / (1.0 + np.asarray(z, dtype=np.float64)))


class des_y1_ia_bias(container):
    """b(z) = A ((1+z)/(1+z0))^eta (bias.py:42-57)."""
    _family = "des_y1_ia"

    def __call__(self, cosmo, z):
        A, eta, z0 = self.params
        return A * ((1.0 + np.asarray(z, dtype=np.float64)) / (1.0 + z0)) ** eta
