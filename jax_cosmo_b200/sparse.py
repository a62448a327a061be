"""Drop-in for jax_cosmo/sparse.py: linear algebra on the block layout `[ny, nx, n]` (a matrix of ny x nx
diagonal blocks of size n -- what `gaussian_cl_covariance(..., sparse=True)` returns with ny = nx = n_cls,
n = n_ell).  Same function names and conventions as the reference (sparse.py:37-389):

    is_sparse, check_sparse, to_dense            layout helpers (host)
    dot(A, B) / dot(A, B, C)                     every vector / dense / sparse combination (sparse.py:72-137)
    inv, slogdet, det                            per-slice inverse and determinant (sparse.py:296-389)

`dot`, `inv`, `slogdet` and `det` run in the CUDA kernels of csrc/jc_sparse.cu through the C ABI
(`jc_sparse_bmm_f64`, `jc_sparse_inv_f64`); NumPy in -> NumPy out, CUDA tensors in -> CUDA tensors out.
There is no CPU fallback.  The fused likelihood / Fisher kernels (`jax_cosmo_b200.likelihood`) remain the fast
path for SPD covariances.
"""
import ctypes as C

import numpy as np

from jax_cosmo_b200 import _native

__all__ = ["is_sparse", "check_sparse", "to_dense", "dot", "inv", "slogdet", "det"]


def _is_tensor(x):
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except ImportError:  # pragma: no cover
        return False


def _ndim(x):
    return x.dim() if _is_tensor(x) else np.asarray(x).ndim


def is_sparse(sparse):
    """True for the 3-D block layout (sparse.py:37-39)."""
    return _ndim(sparse) == 3


def check_sparse(sparse, square=False):
    """Validate the layout (sparse.py:42-49); returns the array."""
    sparse = sparse if _is_tensor(sparse) else np.asarray(sparse)
    if _ndim(sparse) != 3:
        raise ValueError("Expected 3D array of sparse diagonals.")
    if square and sparse.shape[0] != sparse.shape[1]:
        raise ValueError("Expected a square matrix.")
    return sparse


def to_dense(sparse):
    """[ny, nx, n] -> dense [ny*n, nx*n] (sparse.py:52-68)."""
    sparse = np.asarray(check_sparse(sparse.cpu() if _is_tensor(sparse) else sparse))
    ny, nx, n = sparse.shape
    out = np.zeros((ny, n, nx, n), dtype=sparse.dtype)
    idx = np.arange(n)
    out[:, idx, :, idx] = np.moveaxis(sparse, 2, 0)
    return out.reshape(ny * n, nx * n)


def _to_device(x):
    """-> (contiguous CUDA float64 tensor, was_tensor)."""
    import torch

    if not torch.cuda.is_available():
        raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if _is_tensor(x):
        return x.to(device="cuda", dtype=torch.float64).contiguous(), True
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), device="cuda"), False


def _bmm(A, sA, B, sB, out_shape, sC, I, J, K, L):
    import torch

    out = torch.empty(out_shape, dtype=torch.float64, device=A.device)
    with torch.cuda.device(A.device):
        st = _native.load_library().jc_sparse_bmm_f64(
            A.data_ptr(), sA[0], sA[1], sA[2], B.data_ptr(), sB[0], sB[1], sB[2], out.data_ptr(), sC[0], sC[1], sC[2],
            I, J, K, L, torch.cuda.current_stream(A.device).cuda_stream)
    _native.check(st, "jc_sparse_bmm_f64")
    return out


def _dot2(A, B):
    """A @ B for device tensors with at least one sparse operand; strides are (i, j, l) / (j, k, l) / (i, k, l)."""
    if A.dim() == 3:
        ny, nx, n = A.shape
        sA = (nx * n, n, 1)
        if B.dim() == 3:      # sparse @ sparse -> sparse [ny, nz, n]            (sparse.py:235-263)
            nz = B.shape[1]
            return _bmm(A, sA, B, (nz * n, n, 1), (ny, nz, n), (nz * n, n, 1), ny, nx, nz, n)
        if B.dim() == 1:      # sparse @ vec -> vec [ny*n]                        (sparse.py:141-162)
            return _bmm(A, sA, B, (n, 0, 1), (ny * n,), (n, 0, 1), ny, nx, 1, n)
        m = B.shape[1]        # sparse @ dense [nx*n, m] -> dense [ny*n, m]       (sparse.py:166-184)
        return _bmm(A, sA, B, (n * m, 1, m), (ny * n, m), (n * m, 1, m), ny, nx, m, n)
    ny, nx, n = B.shape
    sB = (nx * n, n, 1)
    if A.dim() == 1:          # vec @ sparse -> vec [nx*n]                        (sparse.py:188-209)
        return _bmm(A, (0, n, 1), B, sB, (nx * n,), (0, n, 1), 1, ny, nx, n)
    a = A.shape[0]            # dense [a, ny*n] @ sparse -> dense [a, nx*n]       (sparse.py:213-231)
    return _bmm(A, (ny * n, n, 1), B, sB, (a, nx * n), (nx * n, n, 1), a, ny, nx, n)


def dot(*args):
    """A @ B with at least one sparse operand, or the bilinear form A @ B @ C with dense A, C and sparse B
    (sparse.py:72-137).  Operand kinds follow the array dimension: 1 vector, 2 dense, 3 sparse; the result is
    dense except sparse @ sparse."""
    if len(args) not in (2, 3):
        raise ValueError("Expected 2 or 3 input arrays but got %d." % len(args))
    dims = [_ndim(a) for a in args]
    shapes = [tuple(a.shape) if _is_tensor(a) else np.asarray(a).shape for a in args]
    if len(args) == 2:
        for name, d in zip("AB", dims):
            if d < 1 or d > 3:
                raise ValueError("%s has invalid dimension %d (expected 1 or 2)." % (name, d))
        if 3 not in dims:
            raise ValueError("dot needs at least one sparse (3D) operand.")
        acols = shapes[0][1] * shapes[0][2] if dims[0] == 3 else shapes[0][-1]
        brows = shapes[1][0] * shapes[1][2] if dims[1] == 3 else shapes[1][0]
        if dims[0] == 3 and dims[1] == 3 and shapes[0][2] != shapes[1][2]:
            raise ValueError("Shapes of A %s and B %s not compatible for dot product." % (shapes[0], shapes[1]))
        if acols != brows:
            raise ValueError("Shapes of A %s and B %s not compatible for dot product." % (shapes[0], shapes[1]))
        (A, ta), (B, tb) = _to_device(args[0]), _to_device(args[1])
        out = _dot2(A, B)
        return out if (ta or tb) else out.cpu().numpy()
    if dims != [2, 3, 2]:
        raise ValueError("Can only handle dense @ sparse @ dense bilinear form.")
    if shapes[0][1] != shapes[1][0] * shapes[1][2] or shapes[1][1] * shapes[1][2] != shapes[2][0]:
        raise ValueError("Shapes of A %s, B %s, C %s not compatible for dot product." % tuple(shapes))
    (X, tx), (S, ts), (Y, ty) = (_to_device(a) for a in args)
    XS = _dot2(X, S)                                   # [a, nx*n]
    a, J = XS.shape
    b = Y.shape[1]
    out = _bmm(XS, (J, 1, 0), Y, (b, 1, 0), (a, b), (b, 1, 0), a, J, b, 1)  # (sparse.py:267-292)
    return out if (tx or ts or ty) else out.cpu().numpy()


def _inv_core(sparse, want_inv):
    import torch

    check_sparse(sparse, square=True)
    S, was_tensor = _to_device(sparse)
    P, _, L = S.shape
    inv_t = torch.empty_like(S) if want_inv else None
    sign = torch.empty(L, dtype=torch.float64, device=S.device)
    logdet = torch.empty(L, dtype=torch.float64, device=S.device)
    scratch = torch.empty((L, P, 2 * P), dtype=torch.float64, device=S.device)
    with torch.cuda.device(S.device):
        st = _native.load_library().jc_sparse_inv_f64(
            S.data_ptr(), P, L, inv_t.data_ptr() if want_inv else None, sign.data_ptr(), logdet.data_ptr(),
            scratch.data_ptr(), torch.cuda.current_stream(S.device).cuda_stream)
    _native.check(st, "jc_sparse_inv_f64")
    return inv_t, sign, logdet, was_tensor


def inv(sparse):
    """Inverse of a square sparse matrix, in the same layout (sparse.py:296-315): the inverse of every
    [n_block x n_block] slice `sparse[:, :, l]`."""
    out, _, _, was_tensor = _inv_core(sparse, True)
    return out if was_tensor else out.cpu().numpy()


def slogdet(sparse):
    """(sign, log|det|) of the full block matrix (sparse.py:335-366) = product / sum over the slices."""
    _, sign, logdet, was_tensor = _inv_core(sparse, False)
    s, ld = sign.prod(), logdet.sum()
    return (s, ld) if was_tensor else (float(s.item()), float(ld.item()))


def det(sparse):
    """Determinant of the full block matrix (sparse.py:370-389)."""
    sign, logdet = slogdet(sparse)
    if _is_tensor(sign):
        return sign * logdet.exp()
    return sign * float(np.exp(logdet))


def _named(kinds, name):
    def fn(*args):
        if [_ndim(a) for a in args] != kinds:
            raise ValueError("%s: expected operands of dimension %s" % (name, kinds))
        return dot(*args)
    fn.__name__ = name
    fn.__doc__ = "sparse.py: %s, one strided per-ell product kernel (jc_sparse_bmm_f64) like dot()." % name
    return fn


# the reference's named special cases of dot() (sparse.py:141-292)
sparse_dot_vec = _named([3, 1], "sparse_dot_vec")
sparse_dot_dense = _named([3, 2], "sparse_dot_dense")
vec_dot_sparse = _named([1, 3], "vec_dot_sparse")
dense_dot_sparse = _named([2, 3], "dense_dot_sparse")
sparse_dot_sparse = _named([3, 3], "sparse_dot_sparse")
dense_dot_sparse_dot_dense = _named([2, 3, 2], "dense_dot_sparse_dot_dense")
