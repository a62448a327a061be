"""The XLA-FFI shim (integration/jc_xla_ffi.cc) -- the boundary `north_star` names -- cannot meet a real jaxlib in this
image (no JAX, no xla/ffi/api/ffi.h).  What can be checked is checked:

  * CPU: the handler file compiles and links against libjc_b200.so with the stand-in header tests/ffi_stub/ (its
    Binding::To static_asserts that every Bind() chain matches its Impl signature); the four XLA handler symbols and the
    test hooks are exported; an invalid plan comes back as an ffi::Error (kInvalidArgument), not a crash.
  * GPU: AngularClImpl / AngularClJvpImpl / VjpImpl / GaussianCovImpl driven through fake ffi::Buffers give bitwise what the
    direct C-ABI calls (the ctypes binding) give.
  * whenever `import jax` succeeds (never here): the real jax.ffi registration + jit / jacfwd / grad of
    integration/jax_binding.py against the same values.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import scenarios as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "jax_cosmo_b200", "libjc_xla_ffi_test.so")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def shim():
    from jax_cosmo_b200 import _native
    _native.load_library()  # builds nothing; raises if libjc_b200.so is missing
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-Wall", "-DJC_FFI_TEST_HOOKS",
           "-I" + os.path.join(ROOT, "tests", "ffi_stub"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "integration", "jc_xla_ffi.cc"), "-L" + os.path.join(ROOT, "jax_cosmo_b200"), "-l:libjc_b200.so",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(OUT)


def test_shim_compiles_links_and_exports(shim):
    for sym in ("JcAngularCl", "JcAngularClJvp", "JcVjp", "JcGaussianCov", "jc_ffi_test_angular_cl", "jc_ffi_test_angular_cl_jvp",
                "jc_ffi_test_vjp", "jc_ffi_test_gaussian_cov"):
        assert hasattr(shim, sym), sym
    shim.jc_ffi_test_error_path.restype = C.c_int
    assert shim.jc_ffi_test_error_path() == 3  # ffi::ErrorCode::kInvalidArgument with a message


@pytest.mark.gpu
def test_shim_handlers_equal_direct_calls(shim, jc, torch_cuda):
    torch = torch_cuda
    from jax_cosmo_b200 import _native
    scn = sc.scenario("ffi", sc.PLANCK15, sc.ELL_CFG2[::12], [sc.sources(5, 2.0), sc.lenses(5, 2.0)])
    plan = _native.get_plan(sc.build_probes(scn, jc), scn["ell"], None, None)
    rows = torch.as_tensor(np.concatenate([sc.cosmo_row(sc.PLANCK15)[None], sc.config5_cosmologies(4)]), device="cuda")
    B, P, L, K = rows.shape[0], plan.P, plan.L, 3
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    i64 = C.c_int64
    # forward
    ws = torch.empty(plan.workspace_bytes(B), dtype=torch.uint8, device="cuda")
    cl = torch.empty((B, P, L), dtype=torch.float64, device="cuda")
    st = shim.jc_ffi_test_angular_cl(stream, vp(rows), i64(B), i64(8), i64(plan._h.value), vp(cl), i64(P), i64(L), vp(ws), i64(ws.numel()))
    assert st == 0
    assert torch.equal(cl, plan.angular_cl_device(rows))
    # forward mode
    tang = torch.zeros((K, 8), dtype=torch.float64, device="cuda")
    tang[torch.arange(K), torch.tensor([0, 4, 6])] = 1.0
    need = C.c_size_t()
    _native.check(_native.load_library().jc_workspace_bytes_jvp(plan._h, B * K, C.byref(need)), "ws")
    ws2 = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    cl2 = torch.empty_like(cl)
    dcl = torch.empty((B, K, P, L), dtype=torch.float64, device="cuda")
    st = shim.jc_ffi_test_angular_cl_jvp(stream, vp(rows), i64(B), i64(8), vp(tang), i64(K), i64(plan._h.value), vp(cl2), vp(dcl),
                                         i64(P), i64(L), vp(ws2), i64(ws2.numel()))
    assert st == 0
    ref_cl, ref_d = plan.angular_cl_jvp_device(rows, tang)
    assert torch.equal(cl2, ref_cl) and torch.equal(dcl, ref_d)
    # reverse mode
    cot = torch.randn((B, P, L), dtype=torch.float64, device="cuda")
    grad = torch.empty((B, K), dtype=torch.float64, device="cuda")
    assert shim.jc_ffi_test_vjp(stream, vp(dcl), i64(B), i64(K), i64(P), i64(L), vp(cot), vp(grad)) == 0
    assert torch.equal(grad, _native.vjp_device(dcl, cot))
    # covariance
    noise = torch.as_tensor(plan.noise(), device="cuda")
    cov = torch.empty((B, P, P, L), dtype=torch.float64, device="cuda")
    shim.jc_ffi_test_gaussian_cov.argtypes = [C.c_void_p, C.c_void_p, i64, i64, i64, C.c_void_p, i64, i64, C.c_double, C.c_void_p]
    assert shim.jc_ffi_test_gaussian_cov(stream, vp(cl), B, P, L, vp(noise), plan.T, plan._h.value, 0.3, vp(cov)) == 0
    assert torch.equal(cov, plan.gaussian_cov_device(cl, f_sky=0.3))


@pytest.mark.gpu
def test_real_jax_ffi_when_available(jc, torch_cuda):
    """Runs only where JAX exists (never in this image): the actual jax.ffi custom call + custom_jvp / custom_vjp rules."""
    jax = pytest.importorskip("jax")
    jax.config.update("jax_enable_x64", True)
    if not os.path.exists(os.path.join(ROOT, "jax_cosmo_b200", "libjc_xla_ffi.so")):
        pytest.skip("libjc_xla_ffi.so not built (needs jax.ffi.include_dir())")
    from integration import jax_binding as jb
    scn = sc.scenario("ffi", sc.PLANCK15, sc.ELL_CFG2[::12], [sc.sources(5, 2.0), sc.lenses(5, 2.0)])
    probes = sc.build_probes(scn, jc)
    cosmo = jc.Planck15()
    ref = jc.cl.angular_cl(cosmo, scn["ell"], probes)
    cl = np.asarray(jax.jit(lambda c: jb.angular_cl(c, scn["ell"], probes))(cosmo))
    assert np.array_equal(cl, ref)
    _, jac = jc.cl.angular_cl_jacobian(cosmo, scn["ell"], probes)
    got = jax.jacfwd(lambda c: jb.angular_cl(c, scn["ell"], probes))(cosmo)
    leaves = jax.tree_util.tree_leaves(got)
    assert np.allclose(np.asarray(leaves[0]), jac[0], rtol=1e-12)
