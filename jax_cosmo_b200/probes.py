"""Probes with the reference's constructors (jax_cosmo/probes.py:132-281)."""
import numpy as np

from jax_cosmo_b200.jax_utils import container

__all__ = ["WeakLensing", "NumberCounts", "weak_lensing_kernel", "density_kernel", "nla_kernel"]


def _radial_kernels(probe, cosmo, z):
    """[n_bins, n_z] radial kernels of one probe at redshifts z, by the path's own setup / lensing-efficiency / tracer
    kernels on a grid plan over the probe's tracers (include/jc_b200.h: jc_grid_plan_create_probes); the weak-lensing
    ell factor is left out.  No CPU fallback."""
    import torch

    from jax_cosmo_b200 import _native

    row = cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)
    zz = np.atleast_1d(np.asarray(z, dtype=np.float64)).reshape(-1)
    pb = _native.build_problem([probe], growth=1 if len(row) == 9 else 0)
    out = np.empty((probe.n_tracers, len(zz)))
    for i0 in range(0, len(zz), 512):  # a grid plan holds <= 512 scale factors
        a = 1.0 / (1.0 + zz[i0:i0 + 512])  # z2a, utils.py:2-4
        plan = _native.GridPlan([1.0], a, problem=pb)
        res = plan.evaluate(torch.as_tensor(row[None], device="cuda:%d" % plan.device), want=("kernels",))
        out[:, i0:i0 + 512] = res["kernels"][0].cpu().numpy()
    return out


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

    def kernel(self, cosmo, z, ell):
        """Radial kernels of all bins, shape (n_bins, n_z) (probes.py:188-208): lensing efficiency x (1+z) chi x
        3 H0^2 Omega_m / 2c, plus the NLA term when IA is enabled, times (1 + m), times the ell factor
        sqrt((l-1) l (l+1) (l+2)) / (l+1/2)^2 (probes.py:73)."""
        ell = np.asarray(ell, dtype=np.float64)
        ell_factor = np.sqrt((ell - 1) * ell * (ell + 1) * (ell + 2)) / (ell + 0.5) ** 2
        return ell_factor * _radial_kernels(self, cosmo, z)

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

    def kernel(self, cosmo, z, ell):
        """Radial kernels n_i(z) b_i(z) H(z) of all bins, shape (n_bins, n_z) (probes.py:77-99, 260-272); no ell
        dependence."""
        return _radial_kernels(self, cosmo, z)

    def noise(self):
        """1 / n_gal per bin (probes.py:274-281)."""
        pzs = self.params[0]
        return 1.0 / np.array([pz.gals_per_steradian for pz in pzs])


def weak_lensing_kernel(cosmo, pzs, z, ell):
    """probes.py:17-75: lensing kernels of the bins `pzs` (extended or delta_nz), shape (n_bins, n_z), ell factor included."""
    return WeakLensing(pzs).kernel(cosmo, z, ell)


def density_kernel(cosmo, pzs, bias, z, ell):
    """probes.py:78-100: number-count kernels n(z) b(z) H(z); one bias object or one per bin."""
    return NumberCounts(pzs, bias).kernel(cosmo, z, ell)


def nla_kernel(cosmo, pzs, bias, z, ell):
    """probes.py:103-129: the intrinsic-alignment term alone.  The tracer kernel K2b adds it to the lensing term in place, so
    it is returned as the difference of the two kernel evaluations (both on the GPU path)."""
    return WeakLensing(pzs, ia_bias=bias).kernel(cosmo, z, ell) - WeakLensing(pzs).kernel(cosmo, z, ell)
