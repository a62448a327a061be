"""Cosmology parameter carrier with the reference's constructor and properties
(jax_cosmo/core.py:11-198).  It only carries numbers to the CUDA path: `to_row()` yields the
[8] float64 row in the reference's tree_flatten order (core.py:99-108) that the C ABI consumes."""
import numpy as np

__all__ = ["Cosmology"]

_FIELDS = ("Omega_c", "Omega_b", "h", "n_s", "sigma8", "Omega_k", "w0", "wa")


class Cosmology:
    def __init__(self, Omega_c, Omega_b, h, n_s, sigma8, Omega_k, w0, wa, gamma=None):
        self._Omega_c = Omega_c
        self._Omega_b = Omega_b
        self._h = h
        self._n_s = n_s
        self._sigma8 = sigma8
        self._Omega_k = Omega_k
        self._w0 = w0
        self._wa = wa
        self._gamma = gamma
        self._flags = {"gamma_growth": gamma is not None}
        self._workspace = {}

    def __str__(self):
        return (
            "Cosmological parameters: \n"
            + "".join("    %-9s %s \n" % (k + ":", getattr(self, k))
                      for k in ("h", "Omega_b", "Omega_c", "Omega_k", "w0", "wa", "n_s", "sigma8"))
        )

    __repr__ = __str__

    # pytree-style flattening, same order as the reference
    def tree_flatten(self):
        params = tuple(getattr(self, "_" + k) for k in _FIELDS)
        if self._flags["gamma_growth"]:
            params += (self._gamma,)
        return (params, self._flags)

    @classmethod
    def tree_unflatten(cls, aux_data, children):
        kw = dict(zip(_FIELDS, children[:8]))
        gamma = children[8] if aux_data.get("gamma_growth") else None
        return cls(gamma=gamma, **kw)

    def to_row(self):
        """[8] float64 row for the C ABI; [9] with the growth index appended for a gamma-growth
        cosmology (core.py:56-60,104-105), which selects JC_GROWTH_GAMMA in the plan."""
        row = [float(getattr(self, "_" + k)) for k in _FIELDS]
        if self._flags["gamma_growth"]:
            row.append(float(self._gamma))
        return np.array(row, dtype=np.float64)

    @property
    def Omega(self):
        return 1.0 - self._Omega_k

    @property
    def Omega_b(self):
        return self._Omega_b

    @property
    def Omega_c(self):
        return self._Omega_c

    @property
    def Omega_m(self):
        return self._Omega_b + self._Omega_c

    @property
    def Omega_de(self):
        return self.Omega - self.Omega_m

    @property
    def Omega_k(self):
        return self._Omega_k

    @property
    def k(self):
        return -int(np.sign(self._Omega_k))

    @property
    def sqrtk(self):
        return np.sqrt(np.abs(self._Omega_k))

    @property
    def h(self):
        return self._h

    @property
    def w0(self):
        return self._w0

    @property
    def wa(self):
        return self._wa

    @property
    def n_s(self):
        return self._n_s

    @property
    def sigma8(self):
        return self._sigma8

    @property
    def gamma(self):
        return self._gamma
