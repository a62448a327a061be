"""Default cosmologies (jax_cosmo/parameters.py:10-20): callables that build a `Cosmology`, any parameter overridable by keyword."""
from jax_cosmo_b200.core import Cosmology

__all__ = ["Planck15"]

# best-fit values of Planck 2015 results XIII (TT,TE,EE+lowP+lensing+ext), as tabulated by the reference
_PLANCK15 = {"Omega_c": 0.2589, "Omega_b": 0.04860, "h": 0.6774, "n_s": 0.9667, "sigma8": 0.8159,
             "Omega_k": 0.0, "w0": -1.0, "wa": 0.0}


def Planck15(**overrides):
    """Planck15() or Planck15(sigma8=0.8, gamma=0.55, ...): keyword overrides like the reference's functools.partial."""
    return Cosmology(**dict(_PLANCK15, **overrides))
