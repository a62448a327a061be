"""jax_cosmo_b200 -- B200-native drop-in for jax_cosmo's angular-power-spectrum hot path.

Same names as the reference package (jax_cosmo/__init__.py:10-20) for everything on the path:
Cosmology / Planck15, probes, redshift, bias, power, transfer, and the `cl` alias of angular_cl.
"""
__version__ = "0.1.0"

import jax_cosmo_b200.angular_cl as angular_cl  # module, as in the reference
import jax_cosmo_b200.angular_cl as cl
import jax_cosmo_b200.autograd as autograd
import jax_cosmo_b200.background as background
import jax_cosmo_b200.bias as bias
import jax_cosmo_b200.likelihood as likelihood
import jax_cosmo_b200.power as power
import jax_cosmo_b200.probes as probes
import jax_cosmo_b200.redshift as redshift
import jax_cosmo_b200.sparse as sparse
import jax_cosmo_b200.transfer as transfer
from jax_cosmo_b200.core import *  # noqa: F401,F403
from jax_cosmo_b200.parameters import *  # noqa: F401,F403
from jax_cosmo_b200.parameters import Planck15  # noqa: F401
