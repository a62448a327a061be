"""Drop-in for the matter power spectrum functions of jax_cosmo/power.py.

`halofit` (takahashi2012 / smith2003, power.py:144-272) and `linear` (power.py:81-83) play two roles, as in the
reference: they are the `nonlinear_fn` options of `angular_cl` (recognised by identity, evaluated inside the CUDA
kernels), and they are callable, `nonlinear_fn(cosmo, k, a, transfer_fn)`.  Stand-alone evaluation runs the path's own
setup and power kernels on a grid plan (include/jc_b200.h: jc_grid_plan_create): same Eisenstein-Hu transfer, sigma8
Romberg normalisation, growth table, halofit sigma(R) tables and quirk root as inside `angular_cl`.

    linear_matter_power(cosmo, k, a=1.0, transfer_fn=Eisenstein_Hu)                          power.py:19-53
    nonlinear_matter_power(cosmo, k, a=1.0, transfer_fn=Eisenstein_Hu, nonlinear_fn=halofit)   power.py:265-272
    primordial_matter_power(cosmo, k)                                                          power.py:14-18
    sigmasqr(cosmo, R, transfer_fn)                                                            power.py:56-78

k [h/Mpc] and a broadcast against each other like NumPy arrays (the reference's semantics); the result is squeezed.
No CPU fallback.
"""
import functools

import numpy as np

from jax_cosmo_b200 import _native
from jax_cosmo_b200 import transfer as tklib

__all__ = ["halofit", "linear", "linear_matter_power", "nonlinear_matter_power", "primordial_matter_power", "sigmasqr"]

_MAX_GRID_POINTS = 50_000_000


def primordial_matter_power(cosmo, k):
    """k^n_s (power.py:14-18); a one-line host expression, not a kernel."""
    return np.asarray(k, dtype=np.float64) ** cosmo.n_s


def _transfer_code(transfer_fn):
    fn, kw = transfer_fn, {}
    while isinstance(fn, functools.partial):
        kw = dict(fn.keywords, **kw)
        fn = fn.func
    if fn is not tklib.Eisenstein_Hu or set(kw) - {"type"}:
        raise NotImplementedError("transfer_fn: only jax_cosmo_b200.transfer.Eisenstein_Hu is on the B200 path")
    ttype = kw.get("type", "eisenhu_osc")
    if ttype not in ("eisenhu_osc", "eisenhu"):
        raise NotImplementedError("Eisenstein_Hu type %r (transfer.py:155)" % (ttype,))
    return _native.JC_TF_EH_OSC if ttype == "eisenhu_osc" else _native.JC_TF_EH_NOWIGGLE


def _evaluate(cosmo, k, a, transfer_fn, nonlinear):
    import torch

    row = cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)
    kb, ab = np.broadcast_arrays(np.atleast_1d(np.asarray(k, dtype=np.float64)), np.atleast_1d(np.asarray(a, dtype=np.float64)))
    uk, ik = np.unique(kb, return_inverse=True)
    ua, ia = np.unique(ab, return_inverse=True)
    if len(uk) * min(len(ua), 512) > _MAX_GRID_POINTS:
        raise NotImplementedError("P(k, a) request of %d x %d distinct points exceeds the grid evaluator" % (len(uk), len(ua)))
    tcode = _transfer_code(transfer_fn)
    out = np.empty(kb.size)
    ik, ia = ik.reshape(-1), ia.reshape(-1)
    for a0 in range(0, len(ua), 512):  # a grid plan holds <= 512 scale factors
        sel = (ia >= a0) & (ia < a0 + 512)
        plan = _native.get_grid_plan(uk, ua[a0:a0 + 512], transfer=tcode, nonlinear=nonlinear, growth=1 if len(row) == 9 else 0)
        pk = plan.evaluate(torch.as_tensor(row[None], device="cuda:%d" % plan.device), want=("pk",))["pk"][0]
        idx = torch.as_tensor((ia[sel] - a0) * len(uk) + ik[sel], device=pk.device)
        out[sel] = pk.reshape(-1)[idx].cpu().numpy()
    return out.reshape(kb.shape).squeeze()


def linear_matter_power(cosmo, k, a=1.0, transfer_fn=tklib.Eisenstein_Hu, **kwargs):
    """Linear matter power spectrum [(Mpc/h)^3] (power.py:19-53)."""
    if kwargs:
        transfer_fn = functools.partial(transfer_fn, **kwargs)
    return _evaluate(cosmo, k, a, transfer_fn, _native.JC_PK_LINEAR)


def linear(cosmo, k, a, transfer_fn):
    """`nonlinear_fn` that applies no non-linear correction (power.py:81-83)."""
    return linear_matter_power(cosmo, k, a, transfer_fn)


def halofit(cosmo, k, a, transfer_fn, prescription="takahashi2012"):
    """Halofit non-linear matter power spectrum (power.py:144-262), prescriptions takahashi2012 and smith2003."""
    if prescription not in ("takahashi2012", "smith2003"):
        raise NotImplementedError("halofit prescription %r (power.py:226,244)" % (prescription,))
    code = _native.JC_PK_HALOFIT if prescription == "takahashi2012" else _native.JC_PK_HALOFIT_SMITH
    return _evaluate(cosmo, k, a, transfer_fn, code)


def nonlinear_matter_power(cosmo, k, a=1.0, transfer_fn=tklib.Eisenstein_Hu, nonlinear_fn=halofit):
    """power.py:265-272: nonlinear_fn(cosmo, k, a, transfer_fn=transfer_fn)."""
    return nonlinear_fn(cosmo, k, a, transfer_fn=transfer_fn)


def sigmasqr(cosmo, R, transfer_fn, kmin=0.0001, kmax=1000.0, ksteps=5, **kwargs):
    """sigma^2(R) of the unnormalised spectrum T(k)^2 k^n_s (power.py:56-78; `ksteps` is unused there as well).  Limits other
    than the reference's defaults raise NotImplementedError (JC_ERR_UNSUPPORTED)."""
    import torch

    if kwargs:
        transfer_fn = functools.partial(transfer_fn, **kwargs)
    row = cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)
    plan = _native.get_grid_plan([1.0], [1.0], transfer=_transfer_code(transfer_fn), nonlinear=_native.JC_PK_LINEAR,
                                 growth=1 if len(row) == 9 else 0)
    dev = "cuda:%d" % plan.device
    R_arr = np.ascontiguousarray(np.atleast_1d(np.asarray(R, dtype=np.float64)).reshape(-1))
    out = plan.sigmasqr(torch.as_tensor(row[None], device=dev), torch.as_tensor(R_arr, device=dev), kmin, kmax)[0].cpu().numpy()
    return float(out[0]) if np.ndim(R) == 0 else out.reshape(np.shape(R))
