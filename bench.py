#!/usr/bin/env python
"""bench.py -- C_ell evaluations/s of the B200 angular_cl path on BASELINE config 5.

Workload (config.workload): 3x2pt with 10 source + 10 lens Smail bins (T=20 tracers, P=210 spectra),
100 ell = logspace(1, log10(3000)), halofit, random wCDM cosmologies from the seeded config-5 box
(SURVEY.md 8d).  A *step* = one pass of the hot path over the rank's batch of cosmologies
(default 8192 per GPU = the 65,536-cosmology target / 8: weak scaling, no data-path collective).

  value      whole-job C_ell evaluations (cosmology x ell x pair) per second, inputs resident in HBM
  e2e        same metric through the drop-in host API (pinned host buffers, H2D + D2H inside)
  roofline   dominant kernel of the step (jc_power_kernel or jc_contract_tma_kernel, whichever took longer)
             against the measured FP64 FMA peak, plus the whole step
  cpu_baseline  the NumPy oracle (a port of the reference, see oracle/) on all host cores, bounded sample

`--impl reference` times the reference's CPU path: the reference is pure Python on JAX and JAX is
not installable in this image, so the timed code is the oracle port
(kind="port") on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm runs one single-threaded NumPy worker per host core.  BLAS/OpenMP pools must be pinned to one thread
# BEFORE NumPy is imported (the workers are forked from this process): unpinned, every worker starts its own
# cpu_count()-wide pool and the arm runs 5-10x slower than the cores allow (round-1 finding; torchrun exports
# OMP_NUM_THREADS=1 itself, which is why the N>1 reference runs were faster than N=1).
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ[_v] = "1"

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "C_ell evals/sec (cosmo x ell x pair)"
UNIT = "C_ell/s"
N_ELL, N_SRC, N_LENS = 100, 10, 10
T = N_SRC + N_LENS
P = T * (T + 1) // 2
A = 513
WORKLOAD = "config5: 3x2pt 10+10 Smail bins (T=20, P=210), 100 ell, halofit, wCDM box seed 20240607"


def v0_slots(L, n_src, Pn, halofit=True):
    """SURVEY 8(d) 'convention v0' FP64 issue slots per cosmology; flops = 2 * slots."""
    setup = 5.6e6 if halofit else 0.13e6
    lens = (3.3e6 if n_src > 0 else 0.0) + 0.132e6 * n_src
    power = L * A * (500 if halofit else 280)
    contract = L * A * Pn
    return dict(setup=setup, lens=lens, finish=0.0, power=float(power), contract=float(contract),
                total=setup + lens + power + contract)


def scenario():
    from oracle import scenarios as sc
    return sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(N_SRC, 1.0), sc.lenses(N_LENS, 1.0)])


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def run(self):
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        while not self._halt.is_set():
            try:
                if self.h is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                else:
                    out = subprocess.run(
                        ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
def _oracle_init():
    try:  # belt and braces: pools that ignore the environment variables
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import cl_oracle as o
    from oracle import scenarios as sc
    scn = scenario()
    _oracle_init.ctx = (o, sc.flatten_spec(scn), np.array(scn["ell"]))


def _oracle_worker(rows):
    o, prob, ell = _oracle_init.ctx
    for row in rows:
        o.angular_cl(row, ell, prob)
    return len(rows)


class CpuArm:
    """The reference's CPU path for the bench workload: the NumPy oracle port (oracle/cl_oracle.py; the reference is
    pure Python on JAX and JAX is not installable here) on one single-threaded process per host core.  One persistent
    pool; every timed pass runs `sample` cosmologies of the config-5 box dealt round-robin to the workers."""

    def __init__(self, sample):
        import multiprocessing as mp
        from oracle import scenarios as sc
        self.cores = os.cpu_count() or 1
        self.sample = max(self.cores, int(sample))
        self.rows = sc.config5_cosmologies(65536)[:self.sample]
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_oracle_init)
        self.pool.map(_oracle_worker, [self.rows[i:i + 1] for i in range(self.cores)])  # imports + caches

    def run(self):
        """-> (C_ell/s, seconds) of one pass over the sample."""
        parts = [self.rows[i::self.cores] for i in range(self.cores)]
        t0 = time.perf_counter()
        self.pool.map(_oracle_worker, parts)
        dt = time.perf_counter() - t0
        return self.sample * P * N_ELL / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, rate):
        return {"value": rate, "unit": UNIT, "cores": self.cores, "kind": "port",
                "per_core": rate / self.cores, "blas_threads_per_worker": 1,
                "sample": "%d cosmologies x 210 x 100 per pass, NumPy oracle port of the reference on %d single-threaded "
                          "processes (JAX not installable: the reference itself cannot run)" % (self.sample, self.cores),
                "footnote": "reference source executed unmodified on the NumPy jax shim: ~5 s per ell per cosmology "
                            "(BASELINE.md section 2 item 3; un-jitted interpreter cost, parity use only)"}


def default_cpu_sample():
    """64 cosmologies per host core and pass: ~25 ms each for the vectorised oracle = ~1.6 s of wall clock, ~25 s of CPU work on a
    16-core box (the bounded sample of the measurement contract); `cpu_baseline` and `--impl reference` use the same size."""
    return 64 * (os.cpu_count() or 1)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this workload (oracle port, all host cores)."""
    if rank != 0:
        return
    arm = CpuArm(args.cpu_sample or default_cpu_sample())
    for _ in range(min(args.warmup, 1)):
        arm.run()
    t_tot = 0.0
    for _ in range(args.steps):
        _, dt = arm.run()
        t_tot += dt
    arm.close()
    value = args.steps * arm.sample * P * N_ELL / t_tot
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "cosmologies_per_step": arm.sample},
            "cpu_baseline": arm.describe(value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip() or None
    except Exception:
        return None


def ncu_current():
    """profiles/ncu_current.json: per-kernel figures read from the committed ncu --set full capture (DRAM bytes and
    executed FP64 / DMMA instruction counts per cosmology), stamped with the hash of the kernel sources they were
    captured from.  Stale (sources changed since) -> None: the bench then prints no traffic / executed-work figure
    instead of an outdated one."""
    path = os.path.join(ROOT, "profiles", "ncu_current.json")
    try:
        rec = json.load(open(path))
    except Exception:
        return None, "profiles/ncu_current.json missing"
    if rec.get("csrc_sha256") != csrc_hash():
        return None, "profiles/ncu_current.json is stale (kernel sources changed since the capture)"
    return rec, "profiles/ncu_current.json (%s)" % rec.get("capture", "?")


def csrc_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "jax_cosmo_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def d2h_probe(torch, dev, nbytes=1 << 30, reps=3):
    """Plain pinned cudaMemcpy device->host rate on this box (GB/s): the ceiling of the e2e leg."""
    src = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    dst = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    dst.copy_(src)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5", choices=["config5", "config3", "config4"])
    ap.add_argument("--cosmologies-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="cosmologies per CPU pass (default 16 per host core, the same for cpu_baseline and --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: time the sharded compute only (no exchange)")
    ap.add_argument("--gather-mode", default="peer", choices=["peer", "peer_sm", "nccl", "collective"])
    ap.add_argument("--push-sms", type=int, default=-1, help="SMs of the pusher kernel (gather mode peer_sm)")
    ap.add_argument("--sub-chunk", type=int, default=0, help="cosmologies per compute chunk (K1..K3) of the gather pipeline")
    ap.add_argument("--push-rows", type=int, default=0, help="cosmologies per contraction launch + NVLink push")
    ap.add_argument("--peak-tflops", type=float, default=0.0,
                    help="use this FP64 peak instead of probing (for runs under ncu)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "config5":
        import bench_workloads
        bench_workloads.run(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist

    import jax_cosmo_b200 as jc
    from jax_cosmo_b200 import _native
    from jax_cosmo_b200.distributed import DEFAULT_PUSH_ROWS, DEFAULT_SUB_CHUNK, ShardedAngularCl
    from oracle import scenarios as sc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    scn = scenario()
    probes = sc.build_probes(scn, jc)
    plan = _native.get_plan(probes, scn["ell"], None, None, device=local)
    B = args.cosmologies_per_gpu or 8192
    box = sc.config5_cosmologies(65536)
    if world * B <= 65536:
        rows_all = box[:world * B]
    else:
        rows_all = np.concatenate([box] * (-(-world * B // 65536)))[:world * B]
    rows = np.ascontiguousarray(rows_all[rank * B:(rank + 1) * B])
    cos = torch.as_tensor(rows, device=dev)
    out = torch.empty((B, P, N_ELL), dtype=torch.float64, device=dev)
    ws = plan.workspace(B)
    ws_bytes = ws.numel() * 8
    steps, warmup = args.steps, max(args.warmup, 3)

    # FP64 roofline denominator (not in MEASURED_PEAKS.json): DFMA probe, burst and sustained
    if args.peak_tflops > 0:  # profiler runs: skip the probe launches
        peak_burst = peak_sustained = peak_dmma = args.peak_tflops
        peak_src = "--peak-tflops (earlier jc_fp64_peak_tflops measurement on this pool's B200)"
    else:
        peak_burst = _native.fp64_peak_tflops(0, 0.3)
        peak_sustained = _native.fp64_peak_tflops(0, 2.0)
        peak_dmma = _native.fp64_peak_tflops(1, 0.3)
        peak_src = ("measured live: jc_fp64_peak_tflops DFMA probe, 2 s sustained "
                    "(MEASURED_PEAKS.json holds no FP64 figure)")

    # ---- sharded compute, results resident on the owning GPU (no exchange) ----------------------------------------
    for _ in range(warmup):
        plan.angular_cl_device(cos, out=out, workspace=ws)
    barrier()
    plan.profile_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        plan.angular_cl_device(cos, out=out, workspace=ws)
    e1.record()
    barrier()
    ms_compute = max_over_ranks(e0.elapsed_time(e1))
    stage_ms, stage_n = plan.profile_read()
    plan.profile_enable(False)
    evals_per_step = world * B * P * N_ELL
    value_compute = evals_per_step * steps / (ms_compute * 1e-3)
    n_launches = int(sum(stage_n.values()))

    # ---- N > 1: the same step including the path's one exchange, the gather of the shards on every rank ------------
    gather = None
    ms = ms_compute
    if world > 1 and not args.no_gather:
        sub = args.sub_chunk or DEFAULT_SUB_CHUNK
        push = args.push_rows or DEFAULT_PUSH_ROWS
        if args.gather_mode not in ("peer", "peer_sm"):
            sub = args.sub_chunk or push  # the NCCL pipeline exchanges per compute chunk
        from jax_cosmo_b200.distributed import DEFAULT_PUSH_SMS
        push_sms = args.push_sms if args.push_sms >= 0 else DEFAULT_PUSH_SMS
        sh = ShardedAngularCl(world * B, scn["ell"], probes, gather_mode=args.gather_mode, sub_chunk=sub, push_rows=push,
                              push_sms=push_sms)
        rows_dev = torch.as_tensor(np.ascontiguousarray(rows_all), device=dev)
        for _ in range(warmup):
            sh(rows_dev)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for _ in range(steps):
            sh(rows_dev)
        g1.record()
        barrier()
        ms = max_over_ranks(g0.elapsed_time(g1))
        # correctness on the hardware: the gathered buffer equals the resident result (own rows) and a fresh local
        # evaluation of rows another rank computed and pushed, bitwise
        nb = (rank + 1) % world
        probe_rows = torch.as_tensor(np.ascontiguousarray(rows_all[nb * B:nb * B + 64]), device=dev)
        same = bool(torch.equal(sh.full[rank * B:(rank + 1) * B], out)) and bool(
            torch.equal(sh.full[nb * B:nb * B + 64], plan.angular_cl_device(probe_rows)))
        same = max_over_ranks(0.0 if same else 1.0) == 0.0
        # the exchange alone (no compute): NVLink leg by itself
        ms_x = None
        if sh.mode in ("peer", "peer_sm"):
            for _ in range(2):
                sh._peer.push(rank * B, B)
                sh.barrier()
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for _ in range(steps):
                sh._peer.push(rank * B, B)
                sh.barrier()
            x1.record()
            barrier()
            ms_x = max_over_ranks(x0.elapsed_time(x1)) / steps
        bytes_in = (world - 1) * B * P * N_ELL * 8
        n_chunks = -(-B // sub)
        n_push = (-(-B // push) + 1) if sh.mode in ("peer", "peer_sm") else n_chunks
        aborted = sh._peer.pusher_aborted() if sh._peer is not None else False
        sub, push = sh.sub_chunk, (sh.push_rows if sh.mode in ("peer", "peer_sm") else sh.sub_chunk)
        n_chunks = -(-B // sub)
        n_push = (-(-B // push) + 1) if sh.mode in ("peer", "peer_sm") else n_chunks
        gather = {"mode": sh.mode, "sub_chunk": sub, "push_rows": push,
                  "push_sms": sh.push_sms, "pusher_aborted": aborted, "ms_per_step": ms / steps, "ms_per_step_compute_only": ms_compute / steps,
                  "exposed_ms": (ms - ms_compute) / steps, "ratio_vs_compute_only": ms / ms_compute,
                  "bytes_in_per_gpu_per_step": bytes_in, "bytes_out_per_gpu_per_step": bytes_in,
                  "nvlink_in_gbs_overlapped": bytes_in / (ms / steps * 1e-3) / 1e9,
                  "exchange_alone_ms": ms_x, "nvlink_in_gbs_alone": (bytes_in / (ms_x * 1e-3) / 1e9) if ms_x else None,
                  "value_compute_only": value_compute, "bitwise_equal_to_local": same,
                  "lockstep": bool(sh.mode == "peer" and world > 2 and (world * B) % world == 0),
                  "link_time_over_compute_time": (ms_x / (ms_compute / steps)) if ms_x else None,
                  "note": "every rank ends the step holding the full [%d, %d, %d] result; each finished slice of the contraction "
                          "is pushed into every peer's buffer over NVLink peer memory (copy engines, in lockstep across ranks) while "
                          "the following slices compute; closed by a stream-ordered one-element NCCL all-reduce" % (world * B, P, N_ELL)}
        passes = steps * max(-(-B // int(plan.workspace_layout(ws_bytes).chunk)), 1)
        per_pass = n_launches // passes  # kernels per chunk pass of the compute-only loop (one contraction each)
        n_launches = ((per_pass - 1) * n_chunks + n_push) * steps
        gather["copies_per_step"] = n_push * (world - 1)
        sh.close()
    clocks = sampler.stop()
    value = evals_per_step * steps / (ms * 1e-3)

    # ---- end to end through the drop-in host API (pinned host buffers) ----------------------------
    e2e = None
    if not args.no_e2e:
        rows_pin = torch.as_tensor(rows).pin_memory()
        out_pin = torch.empty((B, P, N_ELL), dtype=torch.float64).pin_memory()
        for _ in range(2):
            jc.cl.angular_cl_batch(rows_pin, scn["ell"], probes, out=out_pin)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            jc.cl.angular_cl_batch(rows_pin, scn["ell"], probes, out=out_pin)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        d2h_bytes = int(out_pin.numel() * 8)
        e2e = {"value": evals_per_step * steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(rows.nbytes), "d2h_bytes_per_step": d2h_bytes}
        # every row of the host result against the device-resident result of the same rows
        diff = 0.0
        for r0 in range(0, B, 1024):
            diff = max(diff, float((out_pin[r0:r0 + 1024].to(dev) - out[r0:r0 + 1024]).abs().max().item()))
        e2e["max_abs_diff_vs_device_path"] = diff
        e2e["rows_compared"] = B
        e2e["d2h_gbs"] = d2h_bytes * steps / dt / 1e9
        if world == 1:
            peak_d2h = d2h_probe(torch, dev)
            e2e["d2h_peak_gbs"] = peak_d2h
            e2e["frac_of_d2h_peak"] = e2e["d2h_gbs"] / peak_d2h
            e2e["bound"] = "device->host copy of the [B, 210, 100] f64 result over PCIe, compute hidden behind it"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    slots = v0_slots(N_ELL, N_SRC, P)
    lo = plan.workspace_layout(ws_bytes)
    chunk = min(int(lo.chunk), B)
    kernels = {"setup": "jc_setup_kernel", "lens": "jc_lens_kernel", "finish": "jc_tracer_finish_kernel",
               "power": "jc_power_kernel", "contract": "jc_contract_tma_kernel"}
    dom = max(("power", "contract", "setup", "lens"), key=lambda k: stage_ms[k])  # dominant kernel of the step
    n_launch = max(stage_n[dom], 1)
    launch_ms = stage_ms[dom] / n_launch              # average launch duration (CUDA events on the launch stream)
    cosmo_per_launch = B * steps / n_launch
    achieved = 2.0 * slots[dom] * cosmo_per_launch / (launch_ms * 1e-3) / 1e12
    step_tflops = 2.0 * slots["total"] * B * steps / (ms_compute * 1e-3) / 1e12
    # executed FP64-pipe work and DRAM traffic: read from the committed ncu capture, refused when stale
    ncu, ncu_src = ncu_current()
    pipe = {}
    traffic = None
    if ncu:
        k = ncu["kernels"]
        # pipe_frac = FP64-datapath busy time at 100 % issue (executed DFMA-class warp instructions x 64 flop, DMMA x 512 flop) / elapsed
        for st in ("setup", "lens", "power", "contract"):
            fl = k[st]["fp64_flops_per_cosmology"]
            pipe[st] = fl * B * steps / (stage_ms[st] * 1e-3) / 1e12 / peak_sustained
        tot = sum(k[st]["fp64_flops_per_cosmology"] for st in ("setup", "lens", "finish", "power", "contract") if st in k)
        pipe["step"] = tot * B * steps / (ms_compute * 1e-3) / 1e12 / peak_sustained
        traffic = k[dom]["dram_bytes_per_cosmology"] * cosmo_per_launch
    alg_bytes = (8.0 * A * N_ELL if dom == "power" else 8.0 * (A * N_ELL + A * T + P * N_ELL)) * cosmo_per_launch
    # "tensor": the dominant kernels are arithmetic bound -- the contraction issues FP64 tensor-core MMAs (DMMA.8x8x4), the power
    # kernel DFMAs on the same FP64 datapath; the denominator is that datapath's measured peak, not the bf16 tensor figure
    roofline = {"bound": "tensor", "bound_detail": "fp64", "kernel": kernels[dom], "achieved": achieved, "peak": peak_sustained,
                "unit": "TFLOP/s", "frac": achieved / peak_sustained, "traffic": traffic,
                "traffic_source": ncu_src, "algorithmic_bytes_per_launch": alg_bytes,
                "pipe_frac": pipe.get(dom), "pipe_frac_step": pipe.get("step"),
                "pipe_frac_setup": pipe.get("setup"), "pipe_frac_lens": pipe.get("lens"),
                "pipe_frac_power": pipe.get("power"), "pipe_frac_contract": pipe.get("contract"),
                "pipe_frac_note": "executed FP64-datapath work (ncu instruction counts of the committed capture: DFMA-class warp "
                                  "instructions x 64 flop + DMMA.8x8x4 x 512 flop) / live stage time / measured FP64 peak -- the pipe "
                                  "utilisation; `frac` follows the SURVEY 8(d) v0 flop convention, which books 500 issue slots per P(k) "
                                  "point and therefore is not a utilisation for the power kernel or the whole step",
                "bound_note": "FP64 arithmetic pipe (DFMA and DMMA share one datapath: jc_fp64_peak_tflops mode 2); "
                              "~550 flop/B, HBM and bf16 tensor peaks of MEASURED_PEAKS.json do not bound this path",
                "peak_source": peak_src, "peak_burst": peak_burst, "peak_dmma": peak_dmma,
                "flops_convention": "SURVEY 8(d) v0 (2 x issue slots): power 500 slots/point, contraction 1 slot/(ell,node,pair)",
                "launch_ms": launch_ms, "cosmologies_per_launch": cosmo_per_launch,
                "frac_v0_step": step_tflops / peak_sustained,
                "ms_setup": stage_ms["setup"] / steps, "ms_lens": stage_ms["lens"] / steps, "ms_finish": stage_ms["finish"] / steps,
                "ms_power": stage_ms["power"] / steps, "ms_contract": stage_ms["contract"] / steps,
                "stage_v0_frac": {k: (2.0 * slots[k] * B * steps / (stage_ms[k] * 1e-3) / 1e12 / peak_sustained)
                                  for k in ("setup", "lens", "power", "contract")}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU arm is timed beside the N = 1 run only
        arm = CpuArm(args.cpu_sample or default_cpu_sample())
        rate, dt = arm.run()
        arm.close()
        cpu = arm.describe(rate)
        cpu["wall_s"] = dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "cosmologies_per_gpu": B, "global_cosmologies": world * B, "chunk": chunk,
                       "l2": "per-step working set (workspace %.1f GB + output %.1f GB) exceeds L2; no flush needed"
                             % (ws_bytes / 1e9, out.numel() * 8 / 1e9),
                       "parallelism": ("cosmology shards, one rank per GPU; the timed step ends with the gather of all shards on every "
                                       "rank (%s)" % gather["mode"]) if gather else
                                      "cosmology shards, one rank per GPU, results resident on the owning GPU (no exchange)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": n_launches,
            "roofline": roofline, "cpu_baseline": cpu, "git_head": git_head()}
    if gather:
        line["gather"] = gather
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
