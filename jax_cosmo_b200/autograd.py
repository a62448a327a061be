"""Reverse mode for the B200 path: what `jax.grad` / `jax.vjp` through the reference's `angular_cl` give
(README.md:24 of the reference: "jax.grad(likelihood)"), here as a `torch.autograd.Function`.

The cosmology has <= 9 parameters, so the vector-Jacobian product is assembled from the forward-mode
kernels: one tangent pass per active parameter (`jc_angular_cl_jvp_f64`) in the forward call, then
`jc_vjp_f64` contracts the saved Jacobian with the incoming cotangent.  No CPU fallback.

    rows = torch.tensor(cosmo.to_row()[None], device="cuda", requires_grad=True)
    cl = angular_cl(rows, ell, probes)              # [B, n_cls, n_ell], differentiable
    loss = ((cl - data) ** 2).sum(); loss.backward()  # rows.grad [B, 8]
"""
import numpy as np

from jax_cosmo_b200 import _native
from jax_cosmo_b200 import power
from jax_cosmo_b200 import transfer as tklib
from jax_cosmo_b200.angular_cl import _PARAM_INDEX, WCDM_PARAMS, _growth, _rows

__all__ = ["angular_cl", "angular_cl_vjp", "value_and_grad"]


def _columns(params, width):
    cols = []
    for name in params:
        if name not in _PARAM_INDEX or _PARAM_INDEX[name] >= width:
            raise ValueError("unknown parameter %r" % (name,))
        cols.append(_PARAM_INDEX[name])
    return cols


def _make_function():
    import torch

    class AngularCl(torch.autograd.Function):
        @staticmethod
        def forward(ctx, rows, plan, cols):
            # grad mode is off inside forward(): a .contiguous() copy of a non-contiguous view would read
            # requires_grad=False, so decide from the autograd context, not from the tensor
            needs_grad = ctx.needs_input_grad[0]
            rows = rows.contiguous()
            if not needs_grad:
                return plan.angular_cl_device(rows)
            tang = torch.zeros((len(cols), rows.shape[1]), dtype=torch.float64, device=rows.device)
            tang[torch.arange(len(cols)), torch.as_tensor(cols)] = 1.0
            cl, dcl = plan.angular_cl_jvp_device(rows.detach(), tang)
            ctx.save_for_backward(dcl)
            ctx.cols = cols
            ctx.width = rows.shape[1]
            return cl

        @staticmethod
        def backward(ctx, grad_cl):
            (dcl,) = ctx.saved_tensors
            g = _native.vjp_device(dcl, grad_cl.contiguous())  # [B, K]
            grad_rows = torch.zeros((dcl.shape[0], ctx.width), dtype=torch.float64, device=dcl.device)
            grad_rows[:, ctx.cols] = g
            return grad_rows, None, None

    return AngularCl


_FUNCTION = None


def angular_cl(rows, ell, probes, transfer_fn=tklib.Eisenstein_Hu, nonlinear_fn=power.halofit, params=None):
    """Differentiable C_ell: `rows` is a CUDA float64 tensor [B, 8] ([B, 9] with gamma) -> [B, n_cls, n_ell].
    Gradients flow to the columns named in `params` (default: every column except Omega_k, i.e. the 7 wCDM
    parameters of BASELINE config 4, plus gamma when present); other columns get zero gradient."""
    global _FUNCTION
    if _FUNCTION is None:
        _FUNCTION = _make_function()
    width = rows.shape[1]
    if params is None:
        params = WCDM_PARAMS + (("gamma",) if width == 9 else ())
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(rows))
    return _FUNCTION.apply(rows, plan, _columns(params, width))


def angular_cl_vjp(cosmo, ell, probes, cotangent, params=WCDM_PARAMS, transfer_fn=tklib.Eisenstein_Hu,
                   nonlinear_fn=power.halofit):
    """(cl [n_cls, n_ell], grad [n_params]) with grad = sum_{p,l} cotangent[p,l] d cl[p,l] / d theta -- the
    `jax.vjp(angular_cl)` product for one cosmology, NumPy in / out."""
    import torch

    rows = _rows(cosmo)
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(rows))
    dev = "cuda:%d" % plan.device
    cols = _columns(params, rows.shape[1])
    tang = np.zeros((len(cols), rows.shape[1]))
    tang[np.arange(len(cols)), cols] = 1.0
    cl, dcl = plan.angular_cl_jvp_device(torch.as_tensor(rows, device=dev), torch.as_tensor(tang, device=dev))
    cot = torch.as_tensor(np.ascontiguousarray(np.asarray(cotangent, dtype=np.float64)), device=dev)
    if cot.numel() != plan.P * plan.L:
        raise ValueError("cotangent must have %d x %d elements" % (plan.P, plan.L))
    g = _native.vjp_device(dcl, cot)
    return cl[0].cpu().numpy(), g[0].cpu().numpy()


def value_and_grad(fn, cosmo, ell, probes, params=WCDM_PARAMS, transfer_fn=tklib.Eisenstein_Hu,
                   nonlinear_fn=power.halofit):
    """`jax.value_and_grad`-style helper: fn(cl) -> scalar tensor, built from torch ops on the CUDA tensor
    cl [n_cls, n_ell].  Returns (value, grad [n_params]) as Python float / NumPy array."""
    import torch

    r = _rows(cosmo)
    plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=_growth(r))
    rows = torch.tensor(r, device="cuda:%d" % plan.device, requires_grad=True)
    cl = angular_cl(rows, ell, probes, transfer_fn, nonlinear_fn, params=params)
    val = fn(cl[0])
    val.backward()
    return float(val.item()), rows.grad[0, _columns(params, r.shape[1])].cpu().numpy()
