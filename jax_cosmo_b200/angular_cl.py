"""Drop-in for jax_cosmo/angular_cl.py on B200: same function names, argument meaning and output
layouts; the arithmetic runs in the sm_100a CUDA kernels behind include/jc_b200.h.

  angular_cl(cosmo, ell, probes, transfer_fn, nonlinear_fn)      angular_cl.py:49-98   -> [n_cls, n_ell]
  noise_cl(ell, probes)                                           angular_cl.py:101-117 -> [n_cls, n_ell]
  gaussian_cl_covariance(ell, probes, cl_signal, cl_noise, ...)   angular_cl.py:120-163
  gaussian_cl_covariance_and_mean(cosmo, ell, probes, ...)        angular_cl.py:166-196

plus the batch form the reference lacks (a batch of cosmologies is the data-parallel axis):

  angular_cl_batch(cosmo_rows[B,8], ell, probes, ...)  -> [B, n_cls, n_ell]

Host inputs / outputs are NumPy float64 arrays ("jax_enable_x64" semantics).  There is no CPU
fallback: without the CUDA library or a GPU every call raises.
"""
import numpy as np

from jax_cosmo_b200 import _native
from jax_cosmo_b200 import power
from jax_cosmo_b200 import transfer as tklib

__all__ = ["angular_cl", "angular_cl_batch", "angular_cl_jvp", "angular_cl_jacobian", "noise_cl", "gaussian_cl_covariance",
           "gaussian_cl_covariance_and_mean"]


def _get_cl_ordering(probes):
    """Pairs (i<=j), row-major upper triangle (angular_cl.py:15-25)."""
    n_tracers = sum([p.n_tracers for p in probes])
    return [(i, j) for i in range(n_tracers) for j in range(i, n_tracers)]


def _pair_index(i, j, T):
    """Arithmetic form of find_index in _get_cov_blocks_ordering (angular_cl.py:34-38)."""
    if i > j:
        i, j = j, i
    return i * T - (i * (i - 1)) // 2 + (j - i)


def _get_cov_blocks_ordering(probes):
    """Index quadruples of the covariance blocks (angular_cl.py:28-46), O(P^2) arithmetic."""
    T = sum([p.n_tracers for p in probes])
    cl_index = _get_cl_ordering(probes)
    return [(_pair_index(i, m, T), _pair_index(j, n, T), _pair_index(i, n, T), _pair_index(j, m, T))
            for (i, j) in cl_index for (m, n) in cl_index]


def _rows(cosmo):
    if hasattr(cosmo, "to_row"):
        return cosmo.to_row()[None, :]
    rows = np.ascontiguousarray(np.asarray(cosmo, dtype=np.float64))
    if rows.ndim == 1:
        rows = rows[None, :]
    if rows.ndim != 2 or rows.shape[1] not in (8, 9):
        raise ValueError("cosmology rows must have shape [B, 8] (Omega_c, Omega_b, h, n_s, sigma8, Omega_k, w0, wa) "
                         "or [B, 9] with the growth index gamma appended (core.py:104-105)")
    return rows


def _growth(rows):
    """JC_GROWTH_GAMMA for 9-column rows (a Cosmology built with gamma=...), else the growth ODE."""
    return 1 if rows.shape[-1] == 9 else 0


def angular_cl(cosmo, ell, probes, transfer_fn=tklib.Eisenstein_Hu, nonlinear_fn=power.halofit):
    """Angular C_ell of all tracer pairs in the Limber approximation -> [n_cls, n_ell]
    (the reference's actual output layout, angular_cl.py:66,98)."""
    rows = _rows(cosmo)
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(rows))
    return plan.angular_cl_host(rows)[0]


def angular_cl_batch(cosmo_rows, ell, probes, transfer_fn=tklib.Eisenstein_Hu, nonlinear_fn=power.halofit,
                     out=None):
    """Batch of cosmologies [B,8] -> [B, n_cls, n_ell].  CUDA tensors in -> CUDA tensor out
    (stream-ordered, no host sync); host arrays in -> host array out."""
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(cosmo_rows))
    try:
        import torch
        if isinstance(cosmo_rows, torch.Tensor) and cosmo_rows.is_cuda:
            return plan.angular_cl_device(cosmo_rows.contiguous(), out=out)
        if isinstance(cosmo_rows, torch.Tensor):
            return plan.angular_cl_host(cosmo_rows.contiguous(), out=out)
    except ImportError:  # pragma: no cover
        pass
    return plan.angular_cl_host(_rows(cosmo_rows), out=out)


_PARAM_INDEX = {"Omega_c": 0, "Omega_b": 1, "h": 2, "n_s": 3, "sigma8": 4, "Omega_k": 5, "w0": 6, "wa": 7, "gamma": 8}
WCDM_PARAMS = ("Omega_c", "Omega_b", "h", "n_s", "sigma8", "w0", "wa")  # BASELINE config 4


def angular_cl_jvp(cosmo, ell, probes, tangents, transfer_fn=tklib.Eisenstein_Hu, nonlinear_fn=power.halofit):
    """Forward-mode derivatives of angular_cl in one call (what `jax.jvp` / `jax.jacfwd` of the
    reference return, docs/notebooks/jax-cosmo-intro.ipynb:989): `tangents` is [K, 8] directions in
    the cosmology-row parameter space.  Returns (cl [B, n_cls, n_ell], dcl [B, K, n_cls, n_ell]) as
    NumPy arrays (B = 1 for a Cosmology object)."""
    import torch

    rows = _rows(cosmo)
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(rows))
    dev = "cuda:%d" % plan.device
    tang = np.ascontiguousarray(np.atleast_2d(np.asarray(tangents, dtype=np.float64)))
    if tang.shape[1] != rows.shape[1]:
        raise ValueError("tangents must have shape [K, %d]" % rows.shape[1])
    rows = torch.as_tensor(rows, device=dev)
    order = _native.direction_order(tang)  # directions that cannot move the tracer kernels last (they may skip K2)
    cl, dcl = plan.angular_cl_jvp_device(rows, torch.as_tensor(np.ascontiguousarray(tang[order]), device=dev))
    dcl = _native.to_host(dcl)
    inverse = np.empty_like(order)
    inverse[order] = np.arange(len(order))
    return _native.to_host(cl), (dcl if np.array_equal(order, np.arange(len(order))) else np.ascontiguousarray(dcl[:, inverse]))


def angular_cl_jacobian(cosmo, ell, probes, params=WCDM_PARAMS, transfer_fn=tklib.Eisenstein_Hu,
                        nonlinear_fn=power.halofit):
    """(cl [n_cls, n_ell], jac [n_params, n_cls, n_ell]): d cl / d theta for the named parameters,
    default the 7 wCDM parameters of BASELINE config 4.  `jac.reshape(n_params, -1).T` is the
    [n_cls*n_ell, n_params] layout of `jax.jacfwd(mean_fn)` (jax-cosmo-intro.ipynb:1039)."""
    width = _rows(cosmo).shape[1]
    tang = np.zeros((len(params), width))
    for k, name in enumerate(params):
        if name not in _PARAM_INDEX or _PARAM_INDEX[name] >= width:
            raise ValueError("unknown parameter %r" % (name,))
        tang[k, _PARAM_INDEX[name]] = 1.0
    cl, dcl = angular_cl_jvp(cosmo, ell, probes, tang, transfer_fn, nonlinear_fn)
    return cl[0], dcl[0]


def noise_cl(ell, probes):
    """Noise contribution to the auto-spectra (angular_cl.py:101-117) -> [n_cls, n_ell]."""
    n_ell = len(np.atleast_1d(ell))
    noise = np.concatenate([np.atleast_1d(p.noise()) * np.ones(p.n_tracers) for p in probes])
    out = np.zeros((len(_get_cl_ordering(probes)), n_ell))
    for k, (i, j) in enumerate(_get_cl_ordering(probes)):
        if i == j:
            out[k] = noise[i]
    return out


def gaussian_cl_covariance(ell, probes, cl_signal, cl_noise, f_sky=0.25, sparse=True):
    """Gaussian covariance (angular_cl.py:120-163).  sparse=True -> [n_cls, n_cls, n_ell] in the
    jax_cosmo.sparse block layout; sparse=False -> dense [(n_cls n_ell), (n_cls n_ell)]."""
    import torch

    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64))
    cl_signal = np.asarray(cl_signal, dtype=np.float64)
    cl_noise = np.asarray(cl_noise, dtype=np.float64)
    plan = _native.get_plan(probes, ell, None, None)
    P, L = plan.P, plan.L
    if cl_signal.shape != (P, L) or cl_noise.shape != (P, L):
        raise ValueError("cl_signal / cl_noise must have shape (%d, %d)" % (P, L))
    # the kernel takes per-tracer noise; recover it from the [P, L] noise_cl layout.  Off-diagonal
    # or ell-dependent noise is folded into the signal so that cl_obs is reproduced exactly.
    pairs = _get_cl_ordering(probes)
    auto = [k for k, (i, j) in enumerate(pairs) if i == j]
    nvec = cl_noise[auto, 0].copy()
    resid = cl_noise.copy()
    resid[auto] -= nvec[:, None]
    dev = "cuda:%d" % plan.device
    cl_dev = torch.as_tensor(cl_signal + resid, device=dev).contiguous()[None]
    cov = plan.gaussian_cov_device(cl_dev, f_sky=f_sky, noise=nvec)[0]
    if sparse:
        return _native.to_host(cov)
    dense = torch.zeros((P, L, P, L), dtype=torch.float64, device=dev)
    dense.diagonal(dim1=1, dim2=3).copy_(cov)  # angular_cl.py:159-162
    return _native.to_host(dense.reshape(P * L, P * L))


def gaussian_cl_covariance_and_mean(cosmo, ell, probes, transfer_fn=tklib.Eisenstein_Hu,
                                    nonlinear_fn=power.halofit, f_sky=0.25, sparse=False):
    """(signal-only flattened mean [n_cls*n_ell], covariance) -- angular_cl.py:166-196."""
    import torch

    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64))
    rows = _rows(cosmo)
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(rows))
    dev = "cuda:%d" % plan.device
    cl_dev = plan.angular_cl_device(torch.as_tensor(rows, device=dev))
    cov = plan.gaussian_cov_device(cl_dev, f_sky=f_sky)[0]
    P, L = plan.P, plan.L
    mean = cl_dev[0].reshape(-1).cpu().numpy()
    if sparse:
        return mean, _native.to_host(cov)
    dense = torch.zeros((P, L, P, L), dtype=torch.float64, device=dev)
    dense.diagonal(dim1=1, dim2=3).copy_(cov)
    return mean, _native.to_host(dense.reshape(P * L, P * L))
