"""Power-spectrum selectors (jax_cosmo/power.py).  `nonlinear_fn=halofit` (takahashi2012,
power.py:144-262) and `nonlinear_fn=linear` (power.py:81-83) are recognised by identity by
angular_cl; the evaluation lives in the CUDA kernels (csrc/jc_pipeline.cu)."""

__all__ = ["halofit", "linear"]


def _selector_only(name):
    raise NotImplementedError(
        "%s is a selector for angular_cl(nonlinear_fn=...) on the B200 path; stand-alone P(k) "
        "evaluation is outside the accelerated hot path (no CPU fallback)" % name)


def halofit(cosmo, k, a, transfer_fn, prescription="takahashi2012"):
    _selector_only("halofit")


def linear(cosmo, k, a, transfer_fn):
    _selector_only("linear")
