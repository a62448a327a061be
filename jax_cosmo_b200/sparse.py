"""Layout helper for the block-"diagonal-of-blocks" covariance [n_cls, n_cls, n_ell]
(jax_cosmo/sparse.py:52-68).  Only `to_dense` is needed on this path; the linear algebra on the
sparse layout (inv / slogdet / dot) is SURVEY 8(f) "next"."""
import numpy as np


def to_dense(sparse):
    sparse = np.asarray(sparse)
    ny, nx, n = sparse.shape
    out = np.zeros((ny, n, nx, n), dtype=sparse.dtype)
    idx = np.arange(n)
    out[:, idx, :, idx] = np.moveaxis(sparse, 2, 0)
    return out.reshape(ny * n, nx * n)
