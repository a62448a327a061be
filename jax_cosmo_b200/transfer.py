"""Transfer-function selectors (jax_cosmo/transfer.py).  On the B200 path `transfer_fn` is an
option recognised by identity: the Eisenstein-Hu "eisenhu_osc" fit is evaluated inside the CUDA
kernels (csrc/jc_internal.cuh: jc_eh_transfer)."""

__all__ = ["Eisenstein_Hu"]


def Eisenstein_Hu(cosmo, k, type="eisenhu_osc"):
    raise NotImplementedError(
        "Eisenstein_Hu is a selector for angular_cl(transfer_fn=...) on the B200 path; "
        "stand-alone T(k) evaluation is outside the accelerated hot path (no CPU fallback)")
