"""Redshift distributions with the reference's constructors (jax_cosmo/redshift.py).

These objects describe n(z); evaluation happens in the CUDA plan kernels (csrc/jc_plan.cu: jc_nz_norm_kernel /
jc_nz_node_kernel / jc_nz_lens_kernel).  All four families of the reference (smail_nz, fu_nz, delta_nz, kde_nz) and
systematic_shift are on the path.  Calling a distribution, `nz(z)`, returns the normalised n(z) like the reference
(redshift.py:27-31), evaluated by the same device functions (jc_nz_eval_f64); there is no CPU fallback."""
from jax_cosmo_b200.jax_utils import container

steradian_to_arcmin2 = 11818102.86004228  # redshift.py:10

__all__ = ["smail_nz", "fu_nz", "kde_nz", "delta_nz", "systematic_shift"]


class redshift_distribution(container):
    def __init__(self, *args, gals_per_arcmin2=1.0, zmax=10.0, **kwargs):
        self._norm = None
        self._gals_per_arcmin2 = gals_per_arcmin2
        super(redshift_distribution, self).__init__(*args, zmax=zmax, **kwargs)

    @property
    def zmax(self):
        return self.config["zmax"]

    @property
    def gals_per_arcmin2(self):
        return self._gals_per_arcmin2

    @property
    def gals_per_steradian(self):
        return self._gals_per_arcmin2 * steradian_to_arcmin2

    def __call__(self, z):
        """Normalised n(z) = pz_fn(z) / simps(pz_fn, 0, zmax, 256) (redshift.py:27-31)."""
        from jax_cosmo_b200 import _native
        return _native.nz_eval(self, z)

    def _describe(self):
        """-> (family, params, shifts) for the jc_nz descriptor."""
        raise NotImplementedError(
            "%s is not supported by the B200 angular_cl path (no fallback)" % type(self).__name__)


class smail_nz(redshift_distribution):
    """n(z) = z^a exp(-(z/z0)^b) (redshift.py:61-77)."""

    def _describe(self):
        a, b, z0 = self.params
        return "smail", (float(a), float(b), float(z0)), []


class systematic_shift(redshift_distribution):
    """pz_fn(z) = parent.pz_fn(clip(z - bias, 0)) (redshift.py:159-171); normalised on its own
    [0, zmax] and with its own gals_per_arcmin2 default, exactly like the reference."""

    def _describe(self):
        parent, bias = self.params[:2]
        fam, p, shifts = parent._describe()
        return fam, p, [float(bias)] + shifts


class fu_nz(redshift_distribution):
    """n(z) = (z^a + z^(ab)) / (z^b + c) (Fu et al. 2008; redshift.py:80-105)."""

    def _describe(self):
        a, b, c = self.params
        return "fu", (float(a), float(b), float(c)), []


class delta_nz(redshift_distribution):
    """Single source plane at z0 (redshift.py:108-123); weak lensing without IA only, as in the reference."""

    def __init__(self, *args, **kwargs):
        super(delta_nz, self).__init__(*args, **kwargs)
        self._norm = 1.0

    def _describe(self):
        return "delta", (float(self.params[0]),), []


class kde_nz(redshift_distribution):
    """Gaussian KDE of a catalogue: kde_nz(zcat, weights, bw=...) (redshift.py:126-156)."""

    def _describe(self):
        import numpy as np
        zcat, weight = self.params[:2]
        zcat = np.ascontiguousarray(np.atleast_1d(np.asarray(zcat, dtype=np.float64)))
        w = np.ascontiguousarray(np.atleast_1d(np.asarray(weight, dtype=np.float64)))
        if w.shape != zcat.shape:
            w = np.ascontiguousarray(np.broadcast_to(w, zcat.shape))
        return "kde", (zcat, w, float(self.config["bw"])), []
