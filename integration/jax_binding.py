"""jax.ffi binding of the B200 path (north_star: XLA-FFI custom call) -- the file a jax_cosmo maintainer adds.

EXPERIMENTAL -- not importable in this repository's image: JAX is not installed and cannot be installed (no network), so
this module has never run under JAX.  What CI does check is the C++ side it binds: integration/jc_xla_ffi.cc is compiled
against a stand-in header, linked with libjc_b200.so and its handlers are driven through fake buffers on the GPU
(tests/test_ffi_shim.py, which also runs this module's functions whenever `import jax` succeeds).  The tested binding of the
same C ABI is jax_cosmo_b200/_native.py (ctypes) with PyTorch owning device memory.  With JAX available:

    import jax; jax.config.update("jax_enable_x64", True)
    from integration.jax_binding import angular_cl          # drop-in for jax_cosmo.angular_cl.angular_cl
    cl = jax.jit(lambda c: angular_cl(c, ell, probes))(cosmo)
    jac = jax.jacfwd(lambda c: angular_cl(c, ell, probes))(cosmo)      # forward mode: jc_angular_cl_jvp_f64
    g = jax.grad(lambda c: loss(angular_cl_rev(c, ell, probes)))(cosmo)  # reverse mode: Jacobian passes + jc_vjp_f64

An opaque custom call carries either a custom_jvp or a custom_vjp rule, not both (JAX cannot transpose the FFI JVP), hence
the two entry points; both share the plan and the kernels.
"""
import ctypes
import os

import numpy as np

import jax
import jax.numpy as jnp

from jax_cosmo_b200 import _native

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_HERE, "..", "jax_cosmo_b200", "libjc_xla_ffi.so"))
for _name, _sym in (("jc_angular_cl", "JcAngularCl"), ("jc_angular_cl_jvp", "JcAngularClJvp"), ("jc_vjp", "JcVjp"),
                    ("jc_gaussian_cov", "JcGaussianCov")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")


# The raw jc_plan* is baked into jitted executables as an attribute: a plan must outlive every executable that may still
# call it, so plans used through this module are pinned here for the life of the process (the 8-entry LRU of
# _native.get_plan would otherwise free them under a cached jit).
_PINNED_PLANS = {}


def _leaves(cosmo):
    """Cosmology leaves in tree_flatten order (core.py:99-108).  jax_cosmo_b200.Cosmology is a plain Python object, not a
    registered pytree: flatten it through its own method; a jax_cosmo.Cosmology (registered) goes through jax."""
    if hasattr(cosmo, "tree_flatten"):
        children, _ = cosmo.tree_flatten()
        return list(children)
    return jax.tree_util.tree_leaves(cosmo)


def _plan_for(cosmo_leaves, ell, probes, transfer_fn, nonlinear_fn):
    growth = 1 if len(cosmo_leaves) == 9 else 0
    plan = _native.get_plan(probes, np.asarray(ell), transfer_fn, nonlinear_fn, growth=growth)
    _PINNED_PLANS[plan._h.value] = plan
    return plan


def _call(rows, plan):
    B = rows.shape[0]
    ws = plan.workspace_bytes(B)
    cl, _ = jax.ffi.ffi_call(
        "jc_angular_cl",
        (jax.ShapeDtypeStruct((B, plan.P, plan.L), jnp.float64), jax.ShapeDtypeStruct((ws,), jnp.uint8)),
        vmap_method="expand_dims")(rows, plan=np.int64(plan._h.value))
    return cl


def _call_jvp(rows, tangents, plan):
    B, K = rows.shape[0], tangents.shape[0]
    need = ctypes.c_size_t()
    lib = _native.load_library()
    if B * K <= 512:  # a Jacobian at one cosmology: room for B*K one-direction entries -> one pass
        _native.check(lib.jc_workspace_bytes_jvp(plan._h, B * K, ctypes.byref(need)), "jc_workspace_bytes_jvp")
    else:             # batches: tangent groups in K1 / K2, one reverse sweep in K3
        _native.check(lib.jc_workspace_bytes_jvp_group(plan._h, B, K, ctypes.byref(need)), "jc_workspace_bytes_jvp_group")
    cl, dcl, _ = jax.ffi.ffi_call(
        "jc_angular_cl_jvp",
        (jax.ShapeDtypeStruct((B, plan.P, plan.L), jnp.float64), jax.ShapeDtypeStruct((B, K, plan.P, plan.L), jnp.float64),
         jax.ShapeDtypeStruct((need.value,), jnp.uint8)))(rows, tangents, plan=np.int64(plan._h.value))
    return cl, dcl


def angular_cl(cosmo, ell, probes, transfer_fn=None, nonlinear_fn=None):
    """Same signature and output layout [n_cls, n_ell] as jax_cosmo.angular_cl.angular_cl (angular_cl.py:49-98);
    differentiable in the cosmology leaves in both modes."""
    leaves = _leaves(cosmo)
    plan = _plan_for(leaves, ell, probes, transfer_fn, nonlinear_fn)

    @jax.custom_jvp
    def f(row):
        return _call(row[None, :], plan)[0]

    @f.defjvp
    def f_jvp(primals, tangents):
        (row,), (drow,) = primals, tangents
        cl, dcl = _call_jvp(row[None, :], drow[None, :], plan)
        return cl[0], dcl[0, 0]

    return f(jnp.stack([jnp.asarray(x, dtype=jnp.float64) for x in leaves]))


def angular_cl_rev(cosmo, ell, probes, transfer_fn=None, nonlinear_fn=None):
    """angular_cl with a reverse-mode rule (jax.grad / jax.vjp): the forward pass saves the Jacobian with respect to
    all leaves (one tangent pass per leaf), the backward pass is jc_vjp_f64."""
    leaves = _leaves(cosmo)
    plan = _plan_for(leaves, ell, probes, transfer_fn, nonlinear_fn)
    n = len(leaves)

    @jax.custom_vjp
    def f(row):
        return _call(row[None, :], plan)[0]

    def fwd(row):
        cl, dcl = _call_jvp(row[None, :], jnp.eye(n, dtype=jnp.float64), plan)
        return cl[0], dcl

    def bwd(dcl, cot):
        grad = jax.ffi.ffi_call("jc_vjp", jax.ShapeDtypeStruct((1, n), jnp.float64))(dcl, cot[None])
        return (grad[0],)

    f.defvjp(fwd, bwd)
    return f(jnp.stack([jnp.asarray(x, dtype=jnp.float64) for x in leaves]))
