"""Multi-GPU form of the batch call: a batch of cosmologies is the data-parallel axis.

One process per GPU (`torch.distributed`); rank r owns the contiguous row block
[r*ceil(B/G), (r+1)*ceil(B/G)) of the cosmology array.  There is **no collective on the data path**
(cosmologies are independent; ell / tracer tables are replicated).  The only exchange is the optional
final all-gather of the per-rank `[B/G, P, L]` blocks over NCCL; with `gather=False`
every rank keeps its shard.  Results do not depend on the sharding (no cross-cosmology reduction).
"""
import numpy as np


def shard_bounds(n_rows, world_size, rank):
    """Contiguous block of rank `rank`: [lo, hi).  Blocks have equal size ceil(n/world) except the tail."""
    per = -(-int(n_rows) // int(world_size))
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def angular_cl_sharded(cosmo_rows, ell, probes, transfer_fn=None, nonlinear_fn=None, group=None,
                       gather=False, compute=None):
    """Compute C_ell for the rows owned by this rank.

    cosmo_rows : [B, 8] array ([B, 9] with the growth index gamma), identical on every rank (cheap: 64 B per cosmology).
    gather     : all-gather the blocks so that every rank returns the full [B, P, L] tensor.
    compute    : callable(rows_shard) -> [n, P, L] tensor; defaults to the CUDA path
                 (`angular_cl_batch` on this rank's device).  Injected by the CPU (gloo) tests.
    Returns (cl, (lo, hi)).
    """
    import torch
    import torch.distributed as dist

    from jax_cosmo_b200 import power, transfer
    from jax_cosmo_b200.angular_cl import angular_cl_batch

    rows = np.ascontiguousarray(np.asarray(cosmo_rows, dtype=np.float64))
    if rows.ndim != 2 or rows.shape[1] not in (8, 9):
        raise ValueError("cosmo_rows must have shape [B, 8] (or [B, 9] with gamma)")
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = shard_bounds(len(rows), world, rank)
    if compute is None:
        tf = transfer.Eisenstein_Hu if transfer_fn is None else transfer_fn
        nl = power.halofit if nonlinear_fn is None else nonlinear_fn

        def compute(shard):
            dev = torch.device("cuda", torch.cuda.current_device())
            return angular_cl_batch(torch.as_tensor(shard, device=dev), ell, probes, tf, nl)

    if hi > lo:
        cl = compute(rows[lo:hi])
    else:  # more ranks than rows: learn the trailing shape from a one-row call
        cl = compute(rows[:1])[:0]
    if not gather or world == 1:
        return cl, (lo, hi)
    # equal-size blocks for all_gather_into_tensor: pad the tail rank(s)
    per = -(-len(rows) // world)
    block = torch.zeros((per,) + tuple(cl.shape[1:]), dtype=cl.dtype, device=cl.device)
    block[: hi - lo] = cl
    full = torch.empty((world * per,) + tuple(cl.shape[1:]), dtype=cl.dtype, device=cl.device)
    dist.all_gather_into_tensor(full, block, group=group)
    return full[: len(rows)], (lo, hi)
