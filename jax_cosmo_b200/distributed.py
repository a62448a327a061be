"""Multi-GPU form of the batch call: a batch of cosmologies is the data-parallel axis.

One process per GPU (`torch.distributed`); rank r owns the contiguous row block
[r*ceil(B/G), (r+1)*ceil(B/G)) of the cosmology array.  There is **no collective while K1..K4 run**
(cosmologies are independent; ell / tracer tables are replicated).  The path's one exchange is the final
gather of the per-rank `[B/G, P, L]` blocks; with `gather=False` every rank keeps its shard.
Results do not depend on the sharding (no cross-cosmology reduction) and are bitwise those of one GPU.

Gather modes (`gather_mode`):

  "peer"        every rank owns a full-size `[B, P, L]` buffer mapped into all ranks (CUDA IPC); the rank computes its
                rows straight into the final layout -- K1..K3 on chunks of `sub_chunk` cosmologies, the contraction per
                `push_rows` cosmologies -- and the copy engines push each finished slice into the peers' buffers over
                NVLink while the following slices compute (csrc/jc_gather.cu): no SM is taken from the FP64 kernels.
                With equal shards and more than two ranks the pushes of a slice run in LOCKSTEP on all ranks (a flag
                barrier per slice by stream memory operations): free-running, eight GPUs drift into each other's
                receivers and the ~100 copies per step reach 58 % of the link rate.  Default on CUDA.
  "peer_sm"     the same pipeline with a persistent pusher kernel on `push_sms` reserved SMs storing each finished slice
                to all peers (cp.async.bulk / st.global on the mapped peer pointers, destinations interleaved).  Uniform
                traffic without any cross-rank synchronisation, but an SM moves only ~55 GB/s over NVLink, so ~16 SMs
                are lost to the FP64 kernels; measured slower than lockstep copy engines at 2, 4 and 8 GPUs.
  "nccl"        the same sub-chunk pipeline with one grouped NCCL send/recv per sub-chunk on a side stream, received
                straight into the final layout (the library baseline the peer path is measured against).
  "collective"  one `all_gather` after the compute (any backend; what the CPU / gloo tests exercise).

`ShardedAngularCl` keeps the plan, the buffers and the streams across calls; `angular_cl_sharded` is the one-shot form.
"""
import numpy as np

DEFAULT_SUB_CHUNK = 1184  # compute chunk of K1..K3: 2 x 592 = two full waves of the setup kernel (148 SMs x 4 CTAs)
DEFAULT_PUSH_ROWS = 592   # cosmologies per contraction launch + NVLink push: 4 per persistent contraction CTA
DEFAULT_PUSH_SMS = 16     # SMs of the pusher kernel ("peer_sm" mode)


def shard_bounds(n_rows, world_size, rank):
    """Contiguous block of rank `rank`: [lo, hi).  Blocks have equal size ceil(n/world) except the tail."""
    per = -(-int(n_rows) // int(world_size))
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def _dist_state(group):
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    return (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)


class ShardedAngularCl:
    """angular_cl for a batch of `n_rows` cosmologies sharded over the ranks of `group`, gathered on every rank.

        sh = ShardedAngularCl(len(rows), ell, probes)          # once: plan, symmetric buffers, handle exchange
        full = sh(rows)                                        # [n_rows, P, L] CUDA tensor on this rank's device

    `rows` is the full [n_rows, 8|9] array (host array or CUDA tensor), identical on every rank -- 64 B per cosmology.
    The call is asynchronous on torch's current stream and ends with a stream-ordered cross-rank barrier, so work
    queued behind it on that stream sees every rank's rows.  The returned tensor is the object's own buffer: it is
    overwritten by the next call."""

    def __init__(self, n_rows, ell, probes, transfer_fn=None, nonlinear_fn=None, group=None, gather_mode="auto",
                 sub_chunk=DEFAULT_SUB_CHUNK, growth=0, push_rows=DEFAULT_PUSH_ROWS, push_sms=DEFAULT_PUSH_SMS):
        import torch
        import torch.distributed as dist

        from jax_cosmo_b200 import _native

        self.group = group
        self.world, self.rank = _dist_state(group)
        self.n_rows = int(n_rows)
        self.per = -(-self.n_rows // self.world)
        self.lo, self.hi = shard_bounds(self.n_rows, self.world, self.rank)
        self.sub_chunk = int(sub_chunk) if sub_chunk and sub_chunk > 0 else max(self.per, 1)
        self.push_rows = int(push_rows) if push_rows and push_rows > 0 else 0
        if not torch.cuda.is_available():
            raise _native.JcError("jax_cosmo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.plan = _native.get_plan(probes, ell, transfer_fn, nonlinear_fn, growth=growth)
        self.device = torch.device("cuda", self.plan.device)
        if gather_mode == "auto":
            gather_mode = "peer" if self.world > 1 else "none"
        if self.world == 1:
            gather_mode = "none"
        if gather_mode not in ("peer", "peer_sm", "nccl", "collective", "none"):
            raise ValueError("gather_mode %r" % (gather_mode,))
        self.mode = gather_mode
        self.push_sms = int(push_sms) if gather_mode == "peer_sm" else 0
        if self.push_sms > 0:
            # the pusher kernel holds `push_sms` SMs for the whole step: size the contraction slices and the compute chunks in
            # whole waves of the remaining SMs (4 cosmologies per persistent contraction CTA, 8 per compute chunk)
            avail = max(torch.cuda.get_device_properties(self.device).multi_processor_count - self.push_sms, 8)
            if push_rows == DEFAULT_PUSH_ROWS:
                self.push_rows = 4 * avail
            if sub_chunk == DEFAULT_SUB_CHUNK:
                self.sub_chunk = 8 * avail
        self._flag = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._peer = None
        self._side = None
        shape = (self.world * self.per, self.plan.P, self.plan.L)
        if self.mode in ("peer", "peer_sm"):
            self._peer = _native.PeerGather(self.plan, shape[0], self.rank, self.world, push_sms=self.push_sms)
            handles = [None] * self.world
            dist.all_gather_object(handles, self._peer.handle, group=group)
            self._peer.connect_ipc(handles)
            self.full = self._peer.full
            dist.barrier(group=group)  # every rank has mapped every buffer before anyone pushes
        else:
            self.full = torch.empty(shape, dtype=torch.float64, device=self.device)
            if self.mode == "nccl":
                self._side = torch.cuda.Stream(device=self.device)

    def _rows_dev(self, cosmo_rows):
        import torch

        if isinstance(cosmo_rows, torch.Tensor):
            if cosmo_rows.shape[0] != self.n_rows:
                raise ValueError("expected %d cosmology rows, got %d" % (self.n_rows, cosmo_rows.shape[0]))
            return cosmo_rows[self.lo:self.hi].to(device=self.device, dtype=torch.float64).contiguous()
        rows = np.ascontiguousarray(np.asarray(cosmo_rows, dtype=np.float64))
        if rows.ndim != 2 or rows.shape[0] != self.n_rows:
            raise ValueError("expected [%d, %d] cosmology rows" % (self.n_rows, self.plan.ncp))
        return torch.as_tensor(rows[self.lo:self.hi], device=self.device)

    def barrier(self):
        """Stream-ordered cross-rank barrier: a one-element all-reduce queued on the current stream."""
        import torch.distributed as dist

        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def compute_shard(self, shard_dev):
        """This rank's rows only, into its block of the full buffer (no exchange)."""
        n = shard_dev.shape[0]
        if n:
            self.plan.angular_cl_device(shard_dev, out=self.full[self.lo:self.lo + n])
        return self.full[self.lo:self.lo + n]

    def __call__(self, cosmo_rows):
        import torch
        import torch.distributed as dist

        shard = self._rows_dev(cosmo_rows)
        n = shard.shape[0]
        if self.mode == "none":
            self.compute_shard(shard)
        elif self.mode in ("peer", "peer_sm"):
            self._peer.compute_and_push(shard, self.lo, self.sub_chunk, self.push_rows,
                                        equal_shards=self.n_rows % self.world == 0)
            self.barrier()
        elif self.mode == "collective":
            self.compute_shard(shard)
            blocks = list(self.full.view(self.world, self.per, self.plan.P, self.plan.L).unbind(0))
            dist.all_gather(blocks, blocks[self.rank].clone(), group=self.group)
        else:  # "nccl": grouped send/recv per sub-chunk on a side stream, straight into the final layout
            cur = torch.cuda.current_stream(self.device)
            counts = [shard_bounds(self.n_rows, self.world, r) for r in range(self.world)]
            for c0 in range(0, self.per, self.sub_chunk):
                nc = max(0, min(self.sub_chunk, n - c0))
                if nc:
                    self.plan.angular_cl_device(shard[c0:c0 + nc], out=self.full[self.lo + c0:self.lo + c0 + nc])
                ev = torch.cuda.Event()
                ev.record(cur)
                self._side.wait_event(ev)
                with torch.cuda.stream(self._side):
                    ops = []
                    for i in range(1, self.world):
                        dst, src = (self.rank + i) % self.world, (self.rank - i) % self.world
                        if nc:
                            ops.append(dist.P2POp(dist.isend, self.full[self.lo + c0:self.lo + c0 + nc], dst, group=self.group))
                        slo, shi = counts[src]
                        ns = max(0, min(self.sub_chunk, (shi - slo) - c0))
                        if ns:
                            ops.append(dist.P2POp(dist.irecv, self.full[slo + c0:slo + c0 + ns], src, group=self.group))
                    if ops:
                        for w in dist.batch_isend_irecv(ops):
                            w.wait()
            cur.wait_stream(self._side)
            self.barrier()
        return self.full[:self.n_rows]

    def close(self):
        self.full = None
        if self._peer is not None:
            self._peer.close()
            self._peer = None


def angular_cl_sharded(cosmo_rows, ell, probes, transfer_fn=None, nonlinear_fn=None, group=None,
                       gather=False, compute=None, gather_mode="auto", sub_chunk=DEFAULT_SUB_CHUNK,
                       push_rows=DEFAULT_PUSH_ROWS):
    """Compute C_ell for the rows owned by this rank (one-shot form).

    cosmo_rows : [B, 8] array ([B, 9] with the growth index gamma), identical on every rank (cheap: 64 B per cosmology).
    gather     : exchange the blocks so that every rank returns the full [B, P, L] tensor (`gather_mode`, module doc).
    compute    : callable(rows_shard) -> [n, P, L] tensor replacing the CUDA path; injected by the CPU (gloo) tests of
                 the host-side logic, always gathered with the "collective" mode.
    Returns (cl, (lo, hi)).
    """
    import torch
    import torch.distributed as dist

    from jax_cosmo_b200 import power, transfer

    rows = np.ascontiguousarray(np.asarray(cosmo_rows, dtype=np.float64))
    if rows.ndim != 2 or rows.shape[1] not in (8, 9):
        raise ValueError("cosmo_rows must have shape [B, 8] (or [B, 9] with gamma)")
    world, rank = _dist_state(group)
    lo, hi = shard_bounds(len(rows), world, rank)
    if compute is None:
        tf = transfer.Eisenstein_Hu if transfer_fn is None else transfer_fn
        nl = power.halofit if nonlinear_fn is None else nonlinear_fn
        mode = gather_mode if (gather and world > 1) else "none"
        sh = ShardedAngularCl(len(rows), ell, probes, tf, nl, group=group, gather_mode=mode, sub_chunk=sub_chunk,
                              growth=1 if rows.shape[1] == 9 else 0, push_rows=push_rows)
        full = sh(rows)
        torch.cuda.current_stream(sh.device).synchronize()
        out = (full if mode != "none" else full[lo:hi]).clone()  # the buffer belongs to `sh`
        sh.close()
        return out, (lo, hi)

    if hi > lo:
        cl = compute(rows[lo:hi])
    else:  # more ranks than rows: learn the trailing shape from a one-row call
        cl = compute(rows[:1])[:0]
    if not gather or world == 1:
        return cl, (lo, hi)
    # equal-size blocks, received straight into the final [world * per, ...] layout (no pad copy of the result)
    per = -(-len(rows) // world)
    full = torch.zeros((world * per,) + tuple(cl.shape[1:]), dtype=cl.dtype, device=cl.device)
    full[lo:hi] = cl
    blocks = list(full.view((world, per) + tuple(cl.shape[1:])).unbind(0))
    dist.all_gather(blocks, blocks[rank].clone(), group=group)
    return full[: len(rows)], (lo, hi)
