"""Probes with the reference's constructors (jax_cosmo/probes.py:132-281)."""
import numpy as np

from jax_cosmo_b200.jax_utils import container

__all__ = ["WeakLensing", "NumberCounts"]


class WeakLensing(container):
    """probes.py:132-223.  params = (redshift_bins, multiplicative_bias[, ia_bias]);
    config = {sigma_e, ia_enabled}."""

    def __init__(self, redshift_bins, ia_bias=None, multiplicative_bias=0.0, sigma_e=0.26, **kwargs):
        if ia_bias is None:
            ia_enabled = False
            args = (redshift_bins, multiplicative_bias)
        else:
            ia_enabled = True
            args = (redshift_bins, multiplicative_bias, ia_bias)
        if "ia_enabled" not in kwargs.keys():
            kwargs["ia_enabled"] = ia_enabled
        super(WeakLensing, self).__init__(*args, sigma_e=sigma_e, **kwargs)

    @property
    def n_tracers(self):
        return len(self.params[0])

    @property
    def zmax(self):
        return max([pz.zmax for pz in self.params[0]])

    def noise(self):
        """sigma_e^2 / n_gal per bin (probes.py:210-223)."""
        pzs = self.params[0]
        ngals = np.array([pz.gals_per_steradian for pz in pzs])
        sigma_e = self.config["sigma_e"]
        if isinstance(sigma_e, list):
            sigma_e = np.array([s for s in sigma_e])
        return sigma_e ** 2 / ngals


class NumberCounts(container):
    """probes.py:226-281.  params = (redshift_bins, bias); config = {has_rsd} (stored, unused --
    as in the reference, probes.py:239-242)."""

    def __init__(self, redshift_bins, bias, has_rsd=False, **kwargs):
        super(NumberCounts, self).__init__(redshift_bins, bias, has_rsd=has_rsd, **kwargs)

    @property
    def zmax(self):
        return max([pz.zmax for pz in self.params[0]])

    @property
    def n_tracers(self):
        return len(self.params[0])

    def noise(self):
        """1 / n_gal per bin (probes.py:274-281)."""
        pzs = self.params[0]
        return 1.0 / np.array([pz.gals_per_steradian for pz in pzs])
