"""Forward-mode throughput: 7 wCDM tangents for a batch of cosmologies on the bench tracer set (10+10 bins, 100 ell)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

scn = sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
probes = sc.build_probes(scn, jc)
plan = _native.get_plan(probes, scn["ell"], None, None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows = torch.as_tensor(sc.config5_cosmologies(B), device="cuda")
tang = np.zeros((7, 8))
tang[np.arange(7), [0, 1, 2, 3, 4, 6, 7]] = 1.0
tang = torch.as_tensor(tang, device="cuda")
for _ in range(2):
    plan.angular_cl_jvp_device(rows, tang)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    plan.angular_cl_jvp_device(rows, tang)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print("JVP: %d cosmologies x 7 tangents: %.2f ms  (%.3e dC_ell/s; %.2f x the cost of %d forward passes at 9.76e9 C_ell/s)"
      % (B, dt * 1e3, B * 7 * 210 * 100 / dt, dt / (B * 7 * 210 * 100 / 9.76e9), 7))
