"""Drop-in for the distance / growth functions of jax_cosmo/background.py, evaluated by the same CUDA setup kernel that
feeds `angular_cl` (a "grid plan", include/jc_b200.h: jc_grid_plan_create / jc_grid_eval_f64): the reference's 256-node
chi(a) table and 128-node growth table with its nearest-node interpolation rule, at the caller's scale factors.

    radial_comoving_distance(cosmo, a)      background.py:199-242   [Mpc/h]
    transverse_comoving_distance(cosmo, a)  background.py:297-344   [Mpc/h]
    angular_diameter_distance(cosmo, a)     background.py:347-368   [Mpc/h]
    growth_factor(cosmo, a)                 background.py:371-398   (ODE or gamma parametrisation, D(1) = 1)
    H(cosmo, a), Esqr(cosmo, a)             background.py:93-143    [km/s/(Mpc/h)], dimensionless
    growth_rate(cosmo, a)                   background.py:401-440, 491-512, 551-584   dlnD/dlna
    Omega_m_a, Omega_de_a, dchioverda, w, f_de   background.py:25-196, 270-294   (jc_grid_background_f64)
    a_of_chi(cosmo, chi)                    background.py:245-267   (jc_a_of_chi_f64)

NumPy in, NumPy out (scalar in -> scalar out, like the reference).  No CPU fallback: without a GPU every call raises.
"""
import numpy as np

from jax_cosmo_b200 import _native

__all__ = ["radial_comoving_distance", "transverse_comoving_distance", "angular_diameter_distance", "growth_factor",
           "H", "Esqr", "growth_rate", "Omega_m_a", "Omega_de_a", "dchioverda", "w", "f_de", "a_of_chi"]

_H0 = 100.0  # constants.py:21


def _evaluate(cosmo, a, name):
    import torch

    row = cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)
    a_arr = np.atleast_1d(np.asarray(a, dtype=np.float64))
    flat = a_arr.reshape(-1)
    out = np.empty_like(flat)
    for i0 in range(0, len(flat), 512):  # a grid plan holds <= 512 scale factors
        part = flat[i0:i0 + 512]
        plan = _native.get_grid_plan([1.0], part, nonlinear=_native.JC_PK_LINEAR, growth=1 if len(row) == 9 else 0)
        res = plan.evaluate(torch.as_tensor(row[None], device="cuda:%d" % plan.device), want=(name,))
        out[i0:i0 + 512] = res[name][0].cpu().numpy()
    out = out.reshape(a_arr.shape)
    return float(out[0]) if np.ndim(a) == 0 else out


def radial_comoving_distance(cosmo, a):
    return _evaluate(cosmo, a, "chi")


def transverse_comoving_distance(cosmo, a):
    return _evaluate(cosmo, a, "chi_transverse")


def angular_diameter_distance(cosmo, a):
    ft = transverse_comoving_distance(cosmo, a)
    return np.asarray(a, dtype=np.float64) * ft if np.ndim(a) else float(a) * ft


def growth_factor(cosmo, a):
    return _evaluate(cosmo, a, "growth")


def H(cosmo, a):
    return _evaluate(cosmo, a, "hubble")


def Esqr(cosmo, a):
    h = H(cosmo, a)
    return (h / _H0) ** 2


def _row(cosmo):
    return cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)


def _background(cosmo, a, name):
    import torch

    row = _row(cosmo)
    a_arr = np.atleast_1d(np.asarray(a, dtype=np.float64))
    flat = a_arr.reshape(-1)
    out = np.empty_like(flat)
    field = _native.GridPlan.BG_FIELDS.index(name)
    for i0 in range(0, len(flat), 512):
        part = flat[i0:i0 + 512]
        plan = _native.get_grid_plan([1.0], part, nonlinear=_native.JC_PK_LINEAR, growth=1 if len(row) == 9 else 0)
        aux = plan.background(torch.as_tensor(row[None], device="cuda:%d" % plan.device))
        out[i0:i0 + 512] = aux[0, field].cpu().numpy()
    out = out.reshape(a_arr.shape)
    return float(out[0]) if np.ndim(a) == 0 else out


def growth_rate(cosmo, a):
    return _background(cosmo, a, "growth_rate")


def Omega_m_a(cosmo, a):
    return _background(cosmo, a, "Omega_m_a")


def Omega_de_a(cosmo, a):
    return _background(cosmo, a, "Omega_de_a")


def dchioverda(cosmo, a):
    return _background(cosmo, a, "dchioverda")


def w(cosmo, a):
    return _background(cosmo, a, "w")


def f_de(cosmo, a):
    return _background(cosmo, a, "f_de")


def a_of_chi(cosmo, chi):
    """Always 1-d, like the reference (np.atleast_1d, background.py:266)."""
    import torch

    row = _row(cosmo)
    plan = _native.get_grid_plan([1.0], [1.0], nonlinear=_native.JC_PK_LINEAR, growth=1 if len(row) == 9 else 0)
    dev = "cuda:%d" % plan.device
    chi_arr = np.ascontiguousarray(np.atleast_1d(np.asarray(chi, dtype=np.float64)).reshape(-1))
    return plan.a_of_chi(torch.as_tensor(row[None], device=dev), torch.as_tensor(chi_arr, device=dev))[0].cpu().numpy()
