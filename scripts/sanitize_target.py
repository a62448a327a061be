"""Small target for compute-sanitizer (memcheck / racecheck) over the kernels added in round 2: the forward-mode throughput path
(tangent groups, reverse-sweep K3, dR = 0 contraction), the tensor-core lens kernel, the tabulated power kernel and the
support-range contraction at the bench tracer set.
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

dev = "cuda"
# forward mode: 5 + 5 bins, 12 ell, 75 cosmologies x 7 directions (> 512 entries: throughput path), ragged last chunk
scn = sc.scenario("s4", sc.PLANCK15, sc.ELL_CFG2[::9], [sc.sources(5, 2.0, True), sc.lenses(5, 2.0, True)])
plan = _native.get_plan(sc.build_probes(scn, jc), scn["ell"], None, None)
rows = torch.as_tensor(sc.config5_cosmologies(75), device=dev)
tang = torch.zeros((7, 8), dtype=torch.float64, device=dev)
tang[torch.arange(7), torch.tensor([0, 1, 2, 3, 4, 6, 7])] = 1.0
res = {}
for name, opts in (("adjoint", {"jvp_adjoint": 1, "jvp_group": 4}), ("groups", {"jvp_adjoint": 0, "jvp_group": 4}),
                   ("single", {"jvp_adjoint": 0, "jvp_group": 1})):
    for k, v in opts.items():
        _native.set_option(k, v)
    res[name] = plan.angular_cl_jvp_device(rows, tang)[1]
_native.set_option("jvp_adjoint", 1)
_native.set_option("jvp_group", 4)
torch.cuda.synchronize()
scale = res["single"].abs().amax(dim=3, keepdim=True)
print("adjoint vs single %.2e  groups vs single %.2e" % (float(((res["adjoint"] - res["single"]).abs() / scale).max()),
                                                         float(((res["groups"] - res["single"]).abs() / scale).max())))
# bench tracer set (10 + 10 bins, 100 ell): tabulated K3, TMA contraction with support ranges, both lens kernels; 37 cosmologies
scn5 = sc.scenario("s5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
plan5 = _native.get_plan(sc.build_probes(scn5, jc), scn5["ell"], None, None)
rows5 = torch.as_tensor(np.ascontiguousarray(sc.config5_cosmologies(37)), device=dev)
out = {}
for mode in (0, 1):
    _native.set_option("lens_mma", mode)
    out[mode] = plan5.angular_cl_device(rows5).clone()
_native.set_option("lens_mma", 0)
torch.cuda.synchronize()
print("lens mma vs scalar %.2e" % float(((out[1] - out[0]).abs() / out[0].abs()).max()))
