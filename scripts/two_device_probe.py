"""Debug aid: one process, plans on cuda:0 and cuda:1 (order given on the command line), which workspace tables get written."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

order = [int(x) for x in sys.argv[1:]] or [0, 1]
scn = sc.scenario("d2", sc.PLANCK15, sc.ELL_CFG2[::4], [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
probes = sc.build_probes(scn, jc)
rows = sc.config5_cosmologies(40)
for dev in order:
    with torch.cuda.device(dev):
        plan = _native.get_plan(probes, scn["ell"], None, None, device=dev)
        ws = plan.workspace(40)
        ws.zero_()
        out = plan.angular_cl_device(torch.as_tensor(rows, device="cuda:%d" % dev), workspace=ws)
        torch.cuda.synchronize(dev)
        l = plan.workspace_layout(ws.numel() * 8)
        lo = {k: getattr(l, k) for k, _ in l._fields_}
        w = ws.cpu().numpy()
        print("dev", dev, "out nonzero", int(np.count_nonzero(out.cpu().numpy())), "ws nonzero", int(np.count_nonzero(w)), "of", w.size,
              "first nonzero idx", int(np.flatnonzero(w)[0]) if np.count_nonzero(w) else -1, "last", int(np.flatnonzero(w)[-1]) if np.count_nonzero(w) else -1,
              "layout", lo, flush=True)
