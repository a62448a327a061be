"""World-size-2 gloo tests (CPU) of the cosmology sharding used for multi-GPU runs
(jax_cosmo_b200/distributed.py).  The CUDA compute is replaced by an injected callable (here the
oracle, acting as the checker's compute) so that the host-side logic -- shard bounds, ragged tails,
the final all-gather, sharding invariance -- is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cl_oracle as o
from oracle import scenarios as sc

ELL = [20.0, 200.0, 2000.0]


def _problem():
    scn = sc.scenario("d", sc.PLANCK15, ELL, [sc.nc([sc.smail(2.0, 4.0, 0.5)], sc.bias("constant", 1.2))], "linear")
    return sc.flatten_spec(scn)


def _compute(shard):
    prob = _problem()
    return torch.as_tensor(np.stack([o.angular_cl(r, ELL, prob) for r in shard]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, gather, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jax_cosmo_b200.distributed import angular_cl_sharded
        rows = sc.config5_cosmologies(n_rows)
        cl, (lo, hi) = angular_cl_sharded(rows, ELL, probes=None, gather=gather, compute=_compute)
        np.save(os.path.join(out_dir, "r%d.npy" % rank), cl.numpy())
        np.save(os.path.join(out_dir, "b%d.npy" % rank), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,gather", [(5, True), (4, False), (1, True)])
def test_sharded_two_ranks(tmp_path, n_rows, gather):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, gather, str(tmp_path)), nprocs=world, join=True)
    rows = sc.config5_cosmologies(n_rows)
    serial = _compute(rows).numpy()
    bounds = [tuple(np.load(tmp_path / ("b%d.npy" % r))) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n_rows and bounds[0][1] == bounds[1][0]
    outs = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    if gather:
        for out in outs:  # every rank holds the full, sharding-independent result
            assert out.shape == serial.shape and np.array_equal(out, serial)
    else:
        assert np.array_equal(np.concatenate(outs), serial)
        for out, (lo, hi) in zip(outs, bounds):
            assert out.shape[0] == hi - lo


def test_shard_bounds_cover_and_partition():
    from jax_cosmo_b200.distributed import shard_bounds
    for n in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) == -(-n // world)


def test_unsharded_passthrough():
    from jax_cosmo_b200.distributed import angular_cl_sharded
    rows = sc.config5_cosmologies(3)
    cl, (lo, hi) = angular_cl_sharded(rows, ELL, probes=None, gather=True, compute=_compute)
    assert (lo, hi) == (0, 3) and np.array_equal(cl.numpy(), _compute(rows).numpy())
    with pytest.raises(ValueError):
        angular_cl_sharded(np.zeros((3, 7)), ELL, probes=None, compute=_compute)
