"""Physical constants of the reference (jax_cosmo/constants.py:9-27); the CUDA kernels carry the
same literals (csrc/jc_internal.cuh)."""
c = 299792.458  # km/s
tcmb = 2.726  # K
rh = 2997.92458  # h^{-1} Mpc
eta_nu = 0.68130
H0 = 100.0  # km/s/(h^{-1} Mpc)
C_1 = 5.0 * 1e-14
rhocrit = 2.7750 * 1e11
