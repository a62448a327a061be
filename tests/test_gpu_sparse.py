"""jax_cosmo_b200.sparse on the GPU (csrc/jc_sparse.cu through the C ABI) against the reference's own
known-answer tests (tests/test_sparse.py of the reference, restated line by line) and against dense NumPy
linear algebra on seeded random block matrices at covariance size (210 x 210 x 100)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp(jc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return jc.sparse


def test_to_dense_reference_kat(sp):  # tests/test_sparse.py:8-31
    X = np.array([[[1, 2, 3], [4, 5, 6], [-1, -2, -3]], [[1, 2, 3], [-4, -5, -6], [7, 8, 9]]])
    answer = np.array([[1, 0, 0, 4, 0, 0, -1, 0, 0], [0, 2, 0, 0, 5, 0, 0, -2, 0], [0, 0, 3, 0, 0, 6, 0, 0, -3],
                       [1, 0, 0, -4, 0, 0, 7, 0, 0], [0, 2, 0, 0, -5, 0, 0, 8, 0], [0, 0, 3, 0, 0, -6, 0, 0, 9]])
    assert_array_equal(sp.to_dense(X), answer)
    with pytest.raises(ValueError):
        sp.to_dense([1, 2, 3])
    with pytest.raises(ValueError):
        sp.to_dense(np.ones((2, 3, 4, 5)))


def test_dot_reference_kat(sp):  # tests/test_sparse.py:34-49
    X1 = [[[1.0, 2], [3, 4], [5, 6]], [[4, 5], [6, 7], [8, 9]]]
    X2 = [[[1.0, -2], [3, -4]], [[5, 4], [6, -7]], [[5, 6], [9, 8]]]
    X1d, X2d = sp.to_dense(X1), sp.to_dense(X2)
    v1, v2 = np.arange(6), np.arange(4)
    assert_allclose(X2d @ v2, sp.dot(X2, v2))
    assert_allclose(X1d @ v1, sp.dot(X1, v1))
    assert_allclose(v2 @ X1d, sp.dot(v2, X1))
    assert_allclose(v1 @ X2d, sp.dot(v1, X2))
    assert_allclose(X1d @ X2d, sp.dot(X1, X2d))
    assert_allclose(X1d @ X2d, sp.dot(X1d, X2))
    assert_allclose(X1d @ X2d, sp.to_dense(sp.dot(X1, X2)))
    assert_allclose(X2d @ X1d, sp.to_dense(sp.dot(X2, X1)))
    with pytest.raises(ValueError):
        sp.dot(X1, np.arange(5))
    with pytest.raises(ValueError):
        sp.dot(X1, X2, X1)


def test_bilinear_reference_kat(sp):  # tests/test_sparse.py:52-60
    X1 = [[[1.0, 2], [3, 4], [5, 6]], [[4, 5], [6, 7], [8, 9]]]
    X2 = [[[1.0, -2], [3, -4]], [[5, 4], [6, -7]], [[5, 6], [9, 8]]]
    X1d, X2d = sp.to_dense(X1), sp.to_dense(X2)
    X12, X21 = sp.dot(X2, X1), sp.dot(X1, X2)
    assert_allclose(X1d @ (X2d @ X1d) @ X2d, sp.dot(X1d, X12, X2d))
    assert_allclose(X2d @ (X1d @ X2d) @ X1d, sp.dot(X2d, X21, X1d))


def test_inv_det_reference_kat(sp):  # tests/test_sparse.py:63-88
    X = np.array([[[1.0, 1.0], [1.0, 1.0]], [[1.0, 1.0], [2.0, 2.0]]])
    assert_allclose(sp.inv(X), np.array([[[2.0, 2.0], [-1.0, -1.0]], [[-1.0, -1.0], [1.0, 1.0]]]))
    with pytest.raises(ValueError):
        sp.inv(np.ones((2, 3, 4)))
    Y = np.array([[[1, 2, 3], [4, 5, 6], [-1, 7, -2]], [[1, 2, 3], [-4, -5, -6], [2, -3, 9]],
                  [[7, 8, 9], [5, -4, 6], [-3, -2, -1]]], dtype=np.float64)
    assert -sp.det(-Y) == sp.det(Y)
    assert_allclose(sp.det(Y), np.linalg.det(sp.to_dense(Y)), rtol=1e-12)
    sign, logdet = sp.slogdet(Y)
    s_ref, l_ref = np.linalg.slogdet(sp.to_dense(Y))
    assert sign == s_ref and abs(logdet - l_ref) < 1e-12 * max(1.0, abs(l_ref))


def test_covariance_size_against_numpy(jc, sp):
    """210 x 210 x 100 (the config-3 covariance shape): inverse, slogdet and every product against per-slice
    NumPy; general (non-symmetric) and SPD inputs; CUDA tensors stay on the device."""
    import torch
    rng = np.random.default_rng(5)
    P, L = 210, 100
    G = rng.normal(size=(P, P, L))
    spd = np.einsum("ikl,jkl->ijl", G, G) / P + np.eye(P)[:, :, None]
    gen = G / np.sqrt(P) + 2.0 * np.eye(P)[:, :, None]
    for S in (spd, gen):
        Sl = np.moveaxis(S, 2, 0)                      # [L, P, P]
        ref_inv = np.moveaxis(np.linalg.inv(Sl), 0, 2)
        got = sp.inv(S)
        assert np.max(np.abs(got - ref_inv)) < 1e-11 * np.max(np.abs(ref_inv))
        s_ref, l_ref = np.linalg.slogdet(Sl)
        sign, logdet = sp.slogdet(S)
        assert sign == np.prod(s_ref) and abs(logdet - l_ref.sum()) < 1e-10 * abs(l_ref.sum())
        # S @ inv(S) = identity blocks
        eye = sp.dot(S, got)
        assert np.max(np.abs(eye - np.eye(P)[:, :, None])) < 1e-10
    v = rng.normal(size=P * L)
    D = rng.normal(size=(P * L, 3))
    vl = v.reshape(P, L)
    assert_allclose(sp.dot(gen, v), np.einsum("ijl,jl->il", gen, vl).reshape(-1), rtol=1e-12, atol=1e-12)
    assert_allclose(sp.dot(v, gen), np.einsum("jl,jkl->kl", vl, gen).reshape(-1), rtol=1e-12, atol=1e-12)
    assert_allclose(sp.dot(gen, D), np.einsum("ijl,jlm->ilm", gen, D.reshape(P, L, 3)).reshape(P * L, 3), rtol=1e-12, atol=1e-12)
    assert_allclose(sp.dot(D.T, gen), np.einsum("mjl,jkl->mkl", D.T.reshape(3, P, L), gen).reshape(3, P * L), rtol=1e-12, atol=1e-12)
    # the notebook's Fisher recipe (docs/notebooks/jax-cosmo-intro.ipynb cell 51) equals the fused Fisher kernel
    J = rng.normal(size=(4, P, L))
    F_sparse = sp.dot(J.reshape(4, -1), sp.inv(spd), J.reshape(4, -1).T)
    F_fused = jc.likelihood.fisher_matrix(J, spd)
    assert_allclose(F_sparse, F_fused, rtol=1e-10)
    # device tensors in -> device tensors out
    St = torch.as_tensor(spd, device="cuda")
    it = sp.inv(St)
    assert it.is_cuda and it.shape == St.shape
    sg, ld = sp.slogdet(St)
    assert sg.is_cuda and float(sg) == 1.0 and abs(float(ld) - np.linalg.slogdet(np.moveaxis(spd, 2, 0))[1].sum()) < 1e-8
    assert sp.dot(St, torch.as_tensor(v, device="cuda")).is_cuda
