"""Parity campaign at the bench configuration: the CUDA path against the NumPy oracle (oracle/cl_oracle.py, pinned to the reference by
tests/golden/) on N random cosmologies of the config-5 box plus the 2^7 corners of a shrunk box, all 210 spectra x 100 ell.
Prints per-stage-independent statistics of the relative error |gpu - oracle| / max_l |oracle| per spectrum.

    PYTHONPATH=. python scripts/parity_sweep.py [n_random=512]

Test infrastructure (it runs the oracle); not part of the product path.
"""
import itertools
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import scenarios as sc  # noqa: E402


def scenario():
    return sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])


def _worker(rows):
    from oracle import cl_oracle as o
    scn = scenario()
    prob = sc.flatten_spec(scn)
    ell = np.array(scn["ell"])
    return np.stack([o.angular_cl(r, ell, prob) for r in rows])


def main():
    import torch

    import jax_cosmo_b200 as jc
    from jax_cosmo_b200 import _native

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    box = sc.config5_cosmologies(65536)
    rng = np.random.default_rng(11)
    rows = box[rng.choice(len(box), n, replace=False)]
    lo, hi = box.min(axis=0), box.max(axis=0)
    free = [i for i in range(8) if hi[i] > lo[i]]
    corners = []
    for bits in itertools.product((0, 1), repeat=len(free)):
        r = box[0].copy()
        for b, i in zip(bits, free):
            r[i] = lo[i] if b == 0 else hi[i]
        corners.append(r)
    rows = np.ascontiguousarray(np.concatenate([rows, np.array(corners)]))
    scn = scenario()
    probes = sc.build_probes(scn, jc)
    plan = _native.get_plan(probes, scn["ell"], None, None)
    gpu = plan.angular_cl_device(torch.as_tensor(rows, device="cuda")).cpu().numpy()
    cores = os.cpu_count() or 1
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        ref = np.concatenate(pool.map(_worker, [p for p in np.array_split(rows, 4 * cores) if len(p)]))
    dt = time.time() - t0
    scale = np.abs(ref).max(axis=2, keepdims=True)
    err = np.abs(gpu - ref) / scale
    per_cosmo = err.max(axis=(1, 2))
    worst = int(per_cosmo.argmax())
    res = dict(n_random=n, n_corners=len(corners), free_parameters=len(free), spectra=int(gpu.shape[1]), ell=int(gpu.shape[2]),
               max_rel_err=float(err.max()), median_of_per_cosmology_max=float(np.median(per_cosmo)),
               p99_of_per_cosmology_max=float(np.quantile(per_cosmo, 0.99)), worst_row=[float(v) for v in rows[worst]],
               max_rel_err_corners=float(per_cosmo[n:].max()), oracle_seconds=round(dt, 1), oracle_processes=cores,
               all_finite=bool(np.isfinite(gpu).all()))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
