"""Drop-in for jax_cosmo/likelihood.py on the sparse block covariance layout (BASELINE config 3).

`gaussian_log_likelihood(data, mu, C, include_logdet=True, inverse_method="inverse")` has the reference's
signature and sign convention (likelihood.py:9-61: -0.5 * (r^T C^-1 r - log det C), r = mu - data).  The
sparse layout [n_cls, n_cls, n_ell] (what `gaussian_cl_covariance_and_mean(..., sparse=True)` returns)
runs in the CUDA kernel csrc/jc_loglike.cu: one CTA per ell slice, packed Cholesky in shared memory.
A dense [N, N] covariance is outside the accelerated path and raises (no CPU fallback).
"""
import numpy as np

from jax_cosmo_b200 import _native

__all__ = ["gaussian_log_likelihood", "gaussian_log_likelihood_batch", "fisher_matrix", "gaussian_log_likelihood_grad"]


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
