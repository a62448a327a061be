"""Default cosmologies (jax_cosmo/parameters.py:10-20)."""
from functools import partial

from jax_cosmo_b200.core import Cosmology

# Planck 2015 paper XII Table 4 final column (best fit)
Planck15 = partial(Cosmology, Omega_c=0.2589, Omega_b=0.04860, Omega_k=0.0, h=0.6774,
                   n_s=0.9667, sigma8=0.8159, w0=-1.0, wa=0.0)
