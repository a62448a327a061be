"""Drop-in for jax_cosmo/likelihood.py on the sparse block covariance layout (BASELINE config 3).

`gaussian_log_likelihood(data, mu, C, include_logdet=True, inverse_method="inverse")` has the reference's
signature and sign convention (likelihood.py:9-61: -0.5 * (r^T C^-1 r - log det C), r = mu - data).  The
sparse layout [n_cls, n_cls, n_ell] (what `gaussian_cl_covariance_and_mean(..., sparse=True)` returns)
runs in the CUDA kernel csrc/jc_loglike.cu: one CTA per ell slice, packed Cholesky in shared memory.
A dense [N, N] covariance is outside the accelerated path and raises (no CPU fallback).
"""
import numpy as np

from jax_cosmo_b200 import _native

__all__ = ["gaussian_log_likelihood", "gaussian_log_likelihood_batch", "fisher_matrix", "gaussian_log_likelihood_grad",
           "gaussian_cl_log_likelihood", "gaussian_cl_log_likelihood_and_grad", "gaussian_cl_log_likelihood_hessian",
           "gaussian_log_likelihood_hessian"]


def gaussian_log_likelihood(data, mu, C, include_logdet=True, inverse_method="inverse"):
    """Gaussian log-likelihood of `data` [N] given mean `mu` [N] and sparse covariance C [P, P, L]
    (N = P*L, cls-major).  `inverse_method` is ignored for sparse covariances, as in the reference."""
    import torch

    C = np.asarray(C, dtype=np.float64)
    if C.ndim == 2:
        return _dense_log_likelihood(data, mu, C, include_logdet, inverse_method)
    if C.ndim != 3 or C.shape[0] != C.shape[1]:
        raise ValueError("C must be a dense [N, N] or a sparse [n_cls, n_cls, n_ell] covariance")
    P, _, L = C.shape
    mu = np.asarray(mu, dtype=np.float64).reshape(-1)
    data = np.asarray(data, dtype=np.float64).reshape(-1)
    if mu.shape != (P * L,) or data.shape != (P * L,):
        raise ValueError("data and mu must have %d elements" % (P * L))
    if not torch.cuda.is_available():
        raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    out = _native.gaussian_loglike_device(torch.as_tensor(data, device="cuda"),
                                          torch.as_tensor(mu[None], device="cuda"),
                                          torch.as_tensor(np.ascontiguousarray(C)[None], device="cuda"), include_logdet)
    return float(out[0].item())


def _dense_log_likelihood(data, mu, C, include_logdet, inverse_method):
    """Dense [N, N] covariance (likelihood.py:44-65): not a kernel of this library -- the factorisations are cuSOLVER's through
    torch.linalg on the GPU (library calls off the hot path; the sparse layout above is the accelerated one).  Same conventions
    as the reference: r = mu - data, result -0.5 (r^T C^-1 r - logdet C)."""
    import torch

    if not torch.cuda.is_available():
        raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    r = torch.as_tensor(np.asarray(mu, dtype=np.float64).reshape(-1) - np.asarray(data, dtype=np.float64).reshape(-1), device="cuda")
    if C.shape != (r.numel(), r.numel()):
        raise ValueError("dense covariance must be [%d, %d]" % (r.numel(), r.numel()))
    Cd = torch.as_tensor(np.ascontiguousarray(C), device="cuda")
    if inverse_method == "inverse":
        y = torch.linalg.inv(Cd) @ r
    elif inverse_method == "cholesky":
        y = torch.cholesky_solve(r[:, None], torch.linalg.cholesky(Cd))[:, 0]
    else:
        raise NotImplementedError("inverse_method %r" % (inverse_method,))
    chi2 = torch.dot(r, y)
    if not include_logdet:
        return float((-0.5 * chi2).item())
    return float((-0.5 * (chi2 - torch.linalg.slogdet(Cd)[1])).item())


def fisher_matrix(jac, C):
    """Fisher matrix F = J^T C^-1 J for a sparse covariance C [P, P, L] and a Jacobian `jac`
    [n_params, P, L] (what `angular_cl_jacobian` returns) -- the reference notebook's
    `sparse.dot(dmu.T, sparse.inv(cov), dmu)` (docs/notebooks/jax-cosmo-intro.ipynb cell 51).
    CUDA tensors with a leading batch dimension ([B, K, P, L], [B, P, P, L]) are also accepted and
    stay on the device."""
    import torch

    if isinstance(jac, torch.Tensor) and jac.is_cuda:
        return _native.fisher_device(jac, C)
    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    jac = np.ascontiguousarray(np.asarray(jac, dtype=np.float64))
    if C.ndim != 3 or C.shape[0] != C.shape[1] or jac.ndim != 3 or jac.shape[1:] != (C.shape[0], C.shape[2]):
        raise ValueError("expected jac [n_params, P, L] and a sparse covariance [P, P, L]")
    if not torch.cuda.is_available():
        raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    out = _native.fisher_device(torch.as_tensor(jac[None], device="cuda"), torch.as_tensor(C[None], device="cuda"))
    return out[0].cpu().numpy()


def gaussian_log_likelihood_grad(data, mu, C, jac):
    """Gradient of `gaussian_log_likelihood(data, mu(theta), C)` with respect to theta at fixed covariance,
    -J^T C^-1 (mu - data): what `jax.grad(likelihood)` of the reference's README returns when the covariance
    is precomputed.  `jac` [n_params, P, L] is `angular_cl_jacobian`'s output; the product runs in the Fisher
    kernel with the residual appended as one more right-hand side (n_params <= 15)."""
    import torch

    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    jac = np.ascontiguousarray(np.asarray(jac, dtype=np.float64))
    if C.ndim != 3 or C.shape[0] != C.shape[1] or jac.ndim != 3 or jac.shape[1:] != (C.shape[0], C.shape[2]):
        raise ValueError("expected jac [n_params, P, L] and a sparse covariance [P, P, L]")
    K, P, L = jac.shape
    r = np.asarray(mu, dtype=np.float64).reshape(P, L) - np.asarray(data, dtype=np.float64).reshape(P, L)
    if not torch.cuda.is_available():
        raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    aug = np.concatenate([jac, r[None]], axis=0)
    F = _native.fisher_device(torch.as_tensor(aug[None], device="cuda"), torch.as_tensor(C[None], device="cuda"))
    return -F[0, :K, K].cpu().numpy()


def gaussian_log_likelihood_batch(data, mu, C, include_logdet=True):
    """Device form for batches: CUDA tensors data [N] or [B, N], mu [B, N], C [B, P, P, L] -> [B]
    (stream-ordered; pairs with `Plan.angular_cl_device` + `Plan.gaussian_cov_device`)."""
    return _native.gaussian_loglike_device(data, mu, C, include_logdet)


def _cl_loglike_setup(cosmo, data, ell, probes, transfer_fn, nonlinear_fn):
    import torch

    from jax_cosmo_b200 import power, transfer
    from jax_cosmo_b200.angular_cl import _growth, _rows

    rows = _rows(cosmo)
    tf = transfer.Eisenstein_Hu if transfer_fn is None else transfer_fn
    nl = power.halofit if nonlinear_fn is None else nonlinear_fn
    plan = _native.get_plan(probes, ell, tf, nl, growth=_growth(rows))
    dev = "cuda:%d" % plan.device
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float64)).reshape(-1)
    if data.size != plan.P * plan.L:
        raise ValueError("data must have %d elements (n_cls * n_ell, cls-major)" % (plan.P * plan.L))
    return rows, plan, torch.as_tensor(rows, device=dev), torch.as_tensor(data, device=dev)


def gaussian_cl_log_likelihood(cosmo, data, ell, probes, f_sky=0.25, include_logdet=True, transfer_fn=None,
                               nonlinear_fn=None):
    """The reference's two calls in one device pass (BASELINE config 3):

        mu, cov = gaussian_cl_covariance_and_mean(cosmo, ell, probes, f_sky=f_sky, sparse=True)   # angular_cl.py:166-196
        lnL = gaussian_log_likelihood(data, mu, cov, include_logdet)                              # likelihood.py:9-61

    without forming the covariance (csrc/jc_cl_loglike.cu: the Gaussian covariance of the spectra is S -> C S C on
    symmetric T x T matrices, so every ell slice is a T x T Cholesky instead of a P x P one).  `cosmo` is a Cosmology
    (-> float) or an array of rows [B, 8|9] (-> array [B])."""
    rows, plan, rows_dev, data_dev = _cl_loglike_setup(cosmo, data, ell, probes, transfer_fn, nonlinear_fn)
    out = plan.gaussian_cl_loglike_device(plan.angular_cl_device(rows_dev), data_dev, f_sky, include_logdet).cpu().numpy()
    return float(out[0]) if hasattr(cosmo, "to_row") else out


def gaussian_cl_log_likelihood_and_grad(cosmo, data, ell, probes, params=None, f_sky=0.25, transfer_fn=None,
                                        nonlinear_fn=None):
    """(lnL, d lnL / d theta) of the likelihood above with BOTH the mean and the covariance (and its determinant)
    depending on the cosmology -- what `jax.grad(likelihood)` returns for the reference's README example
    (README.md:17-27).  The likelihood kernel returns the cotangent d lnL / d cl, the forward-mode pipeline the
    Jacobian d cl / d theta (one fused pass), and `jc_vjp_f64` contracts the two.  `params`: names of the cosmology
    parameters (default: the 7 wCDM parameters, plus gamma for a gamma-growth cosmology); returns (float, [n_params])
    for a Cosmology, ([B], [B, n_params]) for rows."""
    import torch

    from jax_cosmo_b200.angular_cl import _PARAM_INDEX, WCDM_PARAMS

    rows, plan, rows_dev, data_dev = _cl_loglike_setup(cosmo, data, ell, probes, transfer_fn, nonlinear_fn)
    width = rows.shape[1]
    if params is None:
        params = WCDM_PARAMS + (("gamma",) if width == 9 else ())
    tang = np.zeros((len(params), width))
    for k, name in enumerate(params):
        if name not in _PARAM_INDEX or _PARAM_INDEX[name] >= width:
            raise ValueError("unknown parameter %r" % (name,))
        tang[k, _PARAM_INDEX[name]] = 1.0
    order = _native.direction_order(tang)  # directions that cannot move the tracer kernels last (they may skip K2)
    cl, dcl = plan.angular_cl_jvp_device(rows_dev, torch.as_tensor(np.ascontiguousarray(tang[order]), device=rows_dev.device))
    lnl, cot = plan.gaussian_cl_loglike_device(cl, data_dev, f_sky, True, want_cotangent=True)
    grad = _native.vjp_device(dcl, cot)
    lnl, grad = lnl.cpu().numpy(), grad.cpu().numpy()
    inverse = np.empty_like(order)
    inverse[order] = np.arange(len(order))
    grad = grad[:, inverse]
    return (float(lnl[0]), grad[0]) if hasattr(cosmo, "to_row") else (lnl, grad)


def gaussian_cl_log_likelihood_hessian(cosmo, data, ell, probes, params=None, f_sky=0.25, rel_step=1e-6, transfer_fn=None,
                                       nonlinear_fn=None):
    """Second derivatives d2 lnL / d theta_i d theta_j of the likelihood above (mean, covariance and log-determinant all functions
    of the cosmology) -- the matrix `jax.hessian(likelihood)` is asked for in the reference's notebook
    (docs/notebooks/jax-cosmo-intro.ipynb:837-843, F = -hessian at the fiducial cosmology).

    Built from the ANALYTIC gradient (forward-mode Jacobian x likelihood cotangent, `gaussian_cl_log_likelihood_and_grad`) by
    central differences: the 2 K displaced cosmologies theta +- h_i e_i run as ONE batch through the CUDA pipeline,
    H[i, :] = (grad(theta + h_i e_i) - grad(theta - h_i e_i)) / (2 h_i) with h_i = rel_step * max(|theta_i|, 0.1), symmetrised.

    What it is and is not.  The reference's program is piecewise smooth in theta: the halofit root (power.py:113) is a linear
    interpolation whose bracket switches from node to node as the cosmology moves (one switch per ~7e-5 in ln sigma8 over the 513
    Limber nodes), and `jax.hessian` differentiates INSIDE the current brackets.  With the default step the difference window
    (2e-6 relative) almost never contains a switch, the analytic gradient is good to ~1e-12, and the quotient reproduces that
    within-bracket second derivative to about six digits; when a switch does fall inside the window, the two halves of the
    quotient straddle a kink and the entry is off by the kink's share (compare two steps to detect it).  A large step
    (rel_step ~ 1e-3) instead averages over many switches and returns the curvature of the underlying smooth function, which
    differs from jax.hessian's by a few per cent.  For the notebook's fixed-covariance likelihood at the fiducial point the
    Hessian is minus the Fisher matrix, which `fisher_matrix` gives exactly.
    Returns (lnL, grad [K], H [K, K]) for a Cosmology or one row."""
    from jax_cosmo_b200.angular_cl import _PARAM_INDEX, WCDM_PARAMS, _rows

    rows = _rows(cosmo)
    if rows.shape[0] != 1:
        raise ValueError("the Hessian is evaluated at one cosmology (a Cosmology or a single row)")
    width = rows.shape[1]
    if params is None:
        params = WCDM_PARAMS + (("gamma",) if width == 9 else ())
    cols = []
    for name in params:
        if name not in _PARAM_INDEX or _PARAM_INDEX[name] >= width:
            raise ValueError("unknown parameter %r" % (name,))
        cols.append(_PARAM_INDEX[name])
    K = len(cols)
    if not rel_step > 0.0:
        raise ValueError("rel_step must be positive")
    h = rel_step * np.maximum(np.abs(rows[0, cols]), 0.1)
    batch = np.repeat(rows, 2 * K + 1, axis=0)  # row 0: theta; rows 1 + 2 i, 2 + 2 i: theta +- h_i e_i
    for i, c in enumerate(cols):
        batch[1 + 2 * i, c] += h[i]
        batch[2 + 2 * i, c] -= h[i]
    lnl, grad = gaussian_cl_log_likelihood_and_grad(batch, data, ell, probes, params=params, f_sky=f_sky, transfer_fn=transfer_fn,
                                                    nonlinear_fn=nonlinear_fn)
    step = batch[1::2, cols][np.arange(K), np.arange(K)] - batch[2::2, cols][np.arange(K), np.arange(K)]  # the representable 2 h_i
    H = (grad[1::2] - grad[2::2]) / step[:, None]
    return float(lnl[0]), grad[0], 0.5 * (H + H.T)


def gaussian_log_likelihood_hessian(cosmo, data, C, ell, probes, params=None, rel_step=1e-6, transfer_fn=None, nonlinear_fn=None):
    """Hessian in the cosmology of the reference notebook's likelihood (docs/notebooks/jax-cosmo-intro.ipynb:754-766 under
    `jax.hessian`, :837-843):

        lnL(theta) = -1/2 (data - mu(theta))^T C^-1 (data - mu(theta))        C fixed (stop_gradient), no log-determinant

    with `C` the sparse [P, P, L] covariance and `mu = angular_cl(theta)`.  The gradient -J(theta)^T C^-1 (mu(theta) - data) is
    analytic (forward-mode Jacobian, per-ell Cholesky kernel jc_fisher_f64 with the residual as one more right-hand side); the
    Hessian is its central difference over ONE batch of 2 K displaced cosmologies, step small against the spacing of the
    interpolation-bracket switches (see gaussian_cl_log_likelihood_hessian).  At the fiducial point (data = mu(theta)) the residual
    term vanishes and the result is minus the Fisher matrix, which `fisher_matrix` gives without any differencing.
    Returns (lnL, grad [K], H [K, K])."""
    import torch

    from jax_cosmo_b200 import power, transfer
    from jax_cosmo_b200.angular_cl import _PARAM_INDEX, WCDM_PARAMS, _growth, _rows

    rows = _rows(cosmo)
    if rows.shape[0] != 1:
        raise ValueError("the Hessian is evaluated at one cosmology (a Cosmology or a single row)")
    width = rows.shape[1]
    if params is None:
        params = WCDM_PARAMS + (("gamma",) if width == 9 else ())
    cols = []
    for name in params:
        if name not in _PARAM_INDEX or _PARAM_INDEX[name] >= width:
            raise ValueError("unknown parameter %r" % (name,))
        cols.append(_PARAM_INDEX[name])
    K = len(cols)
    if K > 15:
        raise ValueError("at most 15 parameters (the Fisher kernel's right-hand sides)")
    if not rel_step > 0.0:
        raise ValueError("rel_step must be positive")
    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    if C.ndim != 3 or C.shape[0] != C.shape[1]:
        raise ValueError("C must be a sparse [n_cls, n_cls, n_ell] covariance")
    P, _, L = C.shape
    data = np.ascontiguousarray(np.asarray(data, dtype=np.float64)).reshape(-1)
    if data.size != P * L:
        raise ValueError("data must have %d elements (n_cls * n_ell, cls-major)" % (P * L))
    tf = transfer.Eisenstein_Hu if transfer_fn is None else transfer_fn
    nl = power.halofit if nonlinear_fn is None else nonlinear_fn
    plan = _native.get_plan(probes, ell, tf, nl, growth=_growth(rows))
    if (plan.P, plan.L) != (P, L):
        raise ValueError("covariance is [%d, %d, %d] but the probes / ell give %d spectra x %d ell" % (P, P, L, plan.P, plan.L))
    dev = "cuda:%d" % plan.device
    h = rel_step * np.maximum(np.abs(rows[0, cols]), 0.1)
    batch = np.repeat(rows, 2 * K + 1, axis=0)  # row 0: theta; rows 1 + 2 i, 2 + 2 i: theta +- h_i e_i
    for i, c in enumerate(cols):
        batch[1 + 2 * i, c] += h[i]
        batch[2 + 2 * i, c] -= h[i]
    B = batch.shape[0]
    tang = np.zeros((K, width))
    tang[np.arange(K), cols] = 1.0
    cl, dcl = plan.angular_cl_jvp_device(torch.as_tensor(batch, device=dev), torch.as_tensor(tang, device=dev))
    resid = cl - torch.as_tensor(data, device=dev).reshape(1, P, L)                       # mu(theta) - data, [B, P, L]
    aug = torch.cat([dcl, resid[:, None]], dim=1)                                         # [B, K + 1, P, L]
    F = _native.fisher_device(aug, torch.as_tensor(C, device=dev)[None].expand(B, P, P, L).contiguous())
    F = F.cpu().numpy()
    grad = -F[:, :K, K]                                                                    # -J^T C^-1 r
    lnl = -0.5 * F[0, K, K]                                                                # -1/2 r^T C^-1 r
    step = batch[1::2, cols][np.arange(K), np.arange(K)] - batch[2::2, cols][np.arange(K), np.arange(K)]
    H = (grad[1::2] - grad[2::2]) / step[:, None]
    return float(lnl), grad[0], 0.5 * (H + H.T)
