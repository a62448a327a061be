"""Drop-in for jax_cosmo/transfer.py: `Eisenstein_Hu(cosmo, k, type="eisenhu_osc")` (transfer.py:10-156).

The function plays two roles, as in the reference: it is the `transfer_fn` option of `angular_cl` / `power.*`
(recognised by identity; `functools.partial(Eisenstein_Hu, type="eisenhu")` selects the no-wiggle fit), and it is
callable: T(k) is evaluated by the device function the setup kernel uses for its own k grids, on a grid plan
(include/jc_b200.h: jc_grid_eval_f64, `transfer` output).  No CPU fallback."""
import numpy as np

from jax_cosmo_b200 import _native

__all__ = ["Eisenstein_Hu"]


def Eisenstein_Hu(cosmo, k, type="eisenhu_osc"):
    """Eisenstein & Hu matter transfer function at k [h/Mpc]; `type` = "eisenhu_osc" (with baryon wiggles) or
    "eisenhu" (no-wiggle fit)."""
    import torch

    if type not in ("eisenhu_osc", "eisenhu"):
        raise NotImplementedError("Eisenstein_Hu type %r (transfer.py:155)" % (type,))
    row = cosmo.to_row() if hasattr(cosmo, "to_row") else np.asarray(cosmo, dtype=np.float64)
    k_arr = np.atleast_1d(np.asarray(k, dtype=np.float64))
    uk, inv = np.unique(k_arr, return_inverse=True)
    plan = _native.get_grid_plan(uk, [1.0], transfer=_native.JC_TF_EH_OSC if type == "eisenhu_osc" else _native.JC_TF_EH_NOWIGGLE,
                                 nonlinear=_native.JC_PK_LINEAR, growth=1 if len(row) == 9 else 0)
    tk = plan.evaluate(torch.as_tensor(row[None], device="cuda:%d" % plan.device), want=("transfer",))["transfer"][0]
    out = tk.cpu().numpy()[inv.reshape(-1)].reshape(k_arr.shape)
    return float(out[0]) if np.ndim(k) == 0 else out
