"""Why does the peer-push pipeline lose NVLink throughput under compute?  torchrun probe (N ranks): the 14-slice push
schedule of one bench step (7 copies per slice at N = 8) on a side stream while the compute stream runs
  none | sleep (spin kernel) | dgemm (cuBLAS FP64) | k123 (setup + lens + finish + power) | k4 (contraction) | full pipeline.
Prints per case: ms per step (max over ranks) and the implied NVLink ingress rate."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from jax_cosmo_b200.distributed import ShardedAngularCl  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B, P, L = 8192, 210, 100
scn = sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
probes = sc.build_probes(scn, jc)
rows_all = sc.config5_cosmologies(65536)[:world * B]
sh = ShardedAngularCl(world * B, scn["ell"], probes, gather_mode="peer")
plan = sh.plan
rows_dev = torch.as_tensor(np.ascontiguousarray(rows_all), device=dev)
shard = rows_dev[rank * B:(rank + 1) * B].contiguous()
lib = _native.load_library()
side = torch.cuda.Stream(device=dev)
ws = plan.workspace(4096)
cl_tmp = torch.empty((4096, P, L), dtype=torch.float64, device=dev)
a = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
bm = torch.randn((4096, 4096), dtype=torch.float64, device=dev)


def stages(mask, n=4096):
    _native.check(lib.jc_debug_stages_f64(plan._h, mask, shard.data_ptr(), n, cl_tmp.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                          torch.cuda.current_stream().cuda_stream), "stages")


def background(kind):
    if kind == "sleep":
        torch.cuda._sleep(int(14e-3 * 1.9e9))
    elif kind == "dgemm":
        for _ in range(3):
            torch.mm(a, bm)
    elif kind == "k123":
        for _ in range(3):
            stages(15)
    elif kind == "k4":
        stages(15)
        for _ in range(5):
            stages(16)


def pushes(n_slices):
    per = B // n_slices
    for j in range(n_slices):
        sh._peer.push(rank * B + j * per, per)


def timed(fn, steps=5):
    for _ in range(2):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


bytes_in = (world - 1) * B * P * L * 8
for kind in ("none", "sleep", "dgemm", "k123", "k4"):
    for n_slices in (1, 14):
        def step():
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                pushes(n_slices)
            background(kind)
            cur.wait_stream(side)
            sh.barrier()
        ms = timed(step)

        def bg_only():
            background(kind)
            sh.barrier()
        ms_bg = timed(bg_only) if kind != "none" else 0.0
        if rank == 0:
            print("background=%-6s slices=%2d  step %.2f ms  background alone %.2f ms  -> %.0f GB/s in if the step were all link time"
                  % (kind, n_slices, ms, ms_bg, bytes_in / (ms * 1e-3) / 1e9), flush=True)
ms = timed(lambda: sh(rows_dev))
if rank == 0:
    print("full pipeline: %.2f ms" % ms, flush=True)
for env in ("1", "4"):
    pass
sh.close()
dist.destroy_process_group()
