"""Single-call latencies of the drop-in API for BASELINE configs 1-4 (host in, host out), median of 20."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import jax_cosmo_b200 as jc  # noqa: E402
from oracle import scenarios as sc  # noqa: E402


def timeit(fn, n=20):
    fn()
    fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


cosmo = jc.Planck15()
s1 = sc.scenario("cfg1", sc.PLANCK15, sc.ELL_CFG1, [sc.sources(4, 6.5)], "linear")
s2 = sc.scenario("cfg2", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(5, 2.0), sc.lenses(5, 2.0)])
s3 = sc.scenario("cfg3", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
p1, p2, p3 = (sc.build_probes(s, jc) for s in (s1, s2, s3))
print("config 1  angular_cl (4 WL bins, 50 ell, linear)             %.3f ms" % timeit(
    lambda: jc.cl.angular_cl(cosmo, s1["ell"], p1, nonlinear_fn=jc.power.linear)))
print("config 2  angular_cl (5+5 bins, 100 ell, halofit)            %.3f ms" % timeit(
    lambda: jc.cl.angular_cl(cosmo, s2["ell"], p2)))
print("config 3  mean + sparse cov (10+10 bins, 100 ell)            %.3f ms" % timeit(
    lambda: jc.cl.gaussian_cl_covariance_and_mean(cosmo, s3["ell"], p3, sparse=True)))
mu, cov = jc.cl.gaussian_cl_covariance_and_mean(cosmo, s3["ell"], p3, sparse=True)
print("config 3  + gaussian_log_likelihood on the sparse cov         %.3f ms" % timeit(
    lambda: jc.likelihood.gaussian_log_likelihood(1.01 * mu, mu, cov)))
print("config 4  angular_cl_jacobian (5+5 bins, 100 ell, 7 params)   %.3f ms" % timeit(
    lambda: jc.cl.angular_cl_jacobian(cosmo, s2["ell"], p2)))
