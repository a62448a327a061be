"""Host-side limits of the multi-GPU `e2e` leg: device->host bandwidth into pinned memory with every rank copying at once,
with the rank's threads (and therefore the first-touch placement of its pinned pages) left where the launcher put them vs
bound to the GPU's NUMA node.  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/numa_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def gpu_numa_node(dev):
    bus = torch.cuda.get_device_properties(dev)
    pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % pci) as f:
            return pci, int(f.read())
    except OSError:
        return pci, -1


def node_cpus(node):
    with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
        out = []
        for part in f.read().strip().split(","):
            lo, _, hi = part.partition("-")
            out += list(range(int(lo), int(hi or lo) + 1))
        return out


def d2h_gbps(dev, n_bytes=1 << 30, reps=5):
    src = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    dist.barrier()
    return reps * n_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    pci, node = gpu_numa_node(local)
    res = dict(rank=rank, pci=pci, numa_node=node, n_nodes=len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]),
               affinity_before=len(os.sched_getaffinity(0)))
    res["d2h_default_GBps"] = d2h_gbps(local)
    if node >= 0:
        os.sched_setaffinity(0, node_cpus(node))
        res["affinity_after"] = len(os.sched_getaffinity(0))
        res["d2h_bound_GBps"] = d2h_gbps(local)
    else:
        dist.barrier(); dist.barrier()
    print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
