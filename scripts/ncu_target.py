"""Small target for `ncu --set full`: the bench tracer set (10+10 bins, 100 ell, halofit) on 592 cosmologies
(= one cosmology per setup-kernel slot, 4 per persistent contraction CTA), three passes.
    ncu --set full --clock-control none --import-source on -k regex:jc_power -s 1 -c 1 -o gpurun_out/x python scripts/ncu_target.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
scn = sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
plan = _native.get_plan(sc.build_probes(scn, jc), scn["ell"], None, None)
rows = torch.as_tensor(np.ascontiguousarray(sc.config5_cosmologies(65536)[:n]), device="cuda")
out = torch.empty((n, plan.P, plan.L), dtype=torch.float64, device="cuda")
for _ in range(3):
    plan.angular_cl_device(rows, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
