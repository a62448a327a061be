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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "C_ell evals/sec (cosmo x ell x pair)"
UNIT = "C_ell/s"
N_ELL, N_SRC, N_LENS = 100, 10, 10
T = N_SRC + N_LENS
P = T * (T + 1) // 2
A = 513


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
def _oracle_worker(args):
    rows, = args
    from oracle import cl_oracle as o
    from oracle import scenarios as sc
    scn = scenario()
    prob = sc.flatten_spec(scn)
    ell = np.array(scn["ell"])
    for row in rows:
        o.angular_cl(row, ell, prob)
    return len(rows)


def cpu_oracle_rate(n_cosmo, cores):
    """Time the oracle port over n_cosmo cosmologies of the bench workload on `cores` processes."""
    import multiprocessing as mp
    from oracle import scenarios as sc
    rows = sc.config5_cosmologies(65536)[:n_cosmo]
    parts = [rows[i::cores] for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_worker, [(p[:1],) for p in parts])  # warm-up (imports, caches)
        t0 = time.perf_counter()
        pool.map(_oracle_worker, [(p,) for p in parts])
        dt = time.perf_counter() - t0
    return n_cosmo * P * N_ELL / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this workload (oracle port, all host cores)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(cores, min(args.cpu_sample, 8 * cores))
    for _ in range(args.warmup):
        cpu_oracle_rate(cores, cores)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_oracle_rate(sample, cores)
        t_tot += dt
        n_tot += sample
    value = n_tot * P * N_ELL / t_tot
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: 3x2pt 10+10 Smail bins, 100 ell, halofit, wCDM box seed 20240607",
                       "cosmologies_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d cosmologies x 210 x 100 per step, NumPy oracle port of the reference "
                                       "(JAX not installable: reference itself cannot run)" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cosmologies-per-gpu", type=int, default=8192)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--peak-tflops", type=float, default=0.0,
                    help="use this FP64 peak instead of probing (for runs under ncu)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import jax_cosmo_b200 as jc
    from jax_cosmo_b200 import _native
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

    scn = scenario()
    probes = sc.build_probes(scn, jc)
    plan = _native.get_plan(probes, scn["ell"], None, None, device=local)
    B = args.cosmologies_per_gpu
    box = sc.config5_cosmologies(65536)
    rows = box[(rank * B) % 65536:][:B] if (rank * B) % 65536 + B <= 65536 else box[:B]
    rows = np.ascontiguousarray(rows)
    cos = torch.as_tensor(rows, device=dev)
    out = torch.empty((B, P, N_ELL), dtype=torch.float64, device=dev)
    ws = plan.workspace(B)
    ws_bytes = ws.numel() * 8

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

    for _ in range(max(args.warmup, 3)):
        plan.angular_cl_device(cos, out=out, workspace=ws)
    barrier()
    plan.profile_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.angular_cl_device(cos, out=out, workspace=ws)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    stage_ms, stage_n = plan.profile_read()
    plan.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals_per_step = world * B * P * N_ELL
    value = evals_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the drop-in host API (pinned host buffers) ----------------------------
    e2e = None
    if not args.no_e2e:
        rows_pin = torch.as_tensor(rows).pin_memory()
        out_pin = torch.empty((B, P, N_ELL), dtype=torch.float64).pin_memory()
        for _ in range(2):
            jc.cl.angular_cl_batch(rows_pin, scn["ell"], probes, out=out_pin)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            jc.cl.angular_cl_batch(rows_pin, scn["ell"], probes, out=out_pin)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": evals_per_step * args.steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(rows.nbytes), "d2h_bytes_per_step": int(out_pin.numel() * 8)}
        chk = float((out_pin[:4].to(dev) - out[:4]).abs().max().item())
        e2e["max_abs_diff_vs_device_path"] = chk

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
    cosmo_per_launch = B * args.steps / n_launch
    achieved = 2.0 * slots[dom] * cosmo_per_launch / (launch_ms * 1e-3) / 1e12
    # FP64 work the compiled kernels actually execute (ncu, profiles/r01_ncu_summary.md): FP64-pipe warp
    # instructions per (ell, node) point x 64 flop for K3; DMMA m8n8k4 count x 512 flop for K4
    sass_flops = {"power": 186.0 * 2 * N_ELL * A, "contract": 45279.0 * 512}
    # DRAM bytes per launch of `chunk` cosmologies, scaled from the ncu --set full captures at 592 cosmologies
    # (profiles/r01_ncu_v14_metrics.csv / r01_ncu_v15_metrics.csv, r01_ncu_summary.md sections 13-14)
    ncu_dram_per_cosmo = {"power": (42.55e6 + 193.52e6) / 592, "contract": (297.02e6 + 77.66e6) / 592}
    step_tflops = 2.0 * slots["total"] * B * args.steps / (ms * 1e-3) / 1e12
    # "tensor": the dominant kernels are arithmetic bound -- the contraction issues FP64 tensor-core MMAs (DMMA.8x8x4), the power
    # kernel DFMAs on the same FP64 datapath; the denominator is that datapath's measured peak, not the bf16 tensor figure
    roofline = {"bound": "tensor", "bound_detail": "fp64", "kernel": kernels[dom], "achieved": achieved, "peak": peak_sustained,
                "unit": "TFLOP/s", "frac": achieved / peak_sustained,
                "traffic": ncu_dram_per_cosmo[dom] * cosmo_per_launch if dom in ncu_dram_per_cosmo else None,
                "traffic_note": "dram__bytes_read+write per launch from profiles/r01_ncu_summary.md, scaled to the launch size; "
                                "algorithmic bytes per launch: %.3e" % (
                                    (8.0 * A * N_ELL if dom == "power" else 8.0 * (A * N_ELL + A * T + P * N_ELL)) * cosmo_per_launch),
                "bound_note": "FP64 arithmetic pipe (DFMA and DMMA share one datapath: jc_fp64_peak_tflops mode 2); "
                              "~550 flop/B, HBM and bf16 tensor peaks of MEASURED_PEAKS.json do not bound this path",
                "peak_source": peak_src, "peak_burst": peak_burst, "peak_dmma": peak_dmma,
                "flops_convention": "SURVEY 8(d) v0 (2 x issue slots): power 500 slots/point, contraction 1 slot/(ell,node,pair)",
                "achieved_sass": (sass_flops[dom] * cosmo_per_launch / (launch_ms * 1e-3) / 1e12) if dom in sass_flops else None,
                "frac_sass": (sass_flops[dom] * cosmo_per_launch / (launch_ms * 1e-3) / 1e12 / peak_sustained) if dom in sass_flops else None,
                "frac_note": "frac follows the SURVEY 8(d) v0 convention and can exceed 1 for jc_power_kernel: v0 books 500 issue "
                             "slots per P(k) point, the compiled kernel needs 186 FP64 instructions (separable power laws, merged "
                             "divisions, table-driven exp/log); frac_sass = executed FP64 flops / peak is the pipe utilisation",
                "launch_ms": launch_ms, "cosmologies_per_launch": cosmo_per_launch,
                "step": {"achieved": step_tflops, "frac": step_tflops / peak_sustained,
                         "slots_per_cosmology": slots["total"]},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_share": {k: v / max(sum(stage_ms.values()), 1e-9) for k, v in stage_ms.items()},
                "stage_v0_frac": {k: (2.0 * slots[k] * B * args.steps / (stage_ms[k] * 1e-3) / 1e12 / peak_sustained)
                                  for k in ("setup", "lens", "power", "contract")}}
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU arm is timed beside the N = 1 run only
        cores = os.cpu_count() or 1
        sample = max(cores, min(args.cpu_sample, 8 * cores))
        rate, dt = cpu_oracle_rate(sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d cosmologies x 210 x 100 (%.1f s wall), NumPy oracle port on %d processes" % (sample, dt, cores)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: 3x2pt 10+10 Smail bins (T=20, P=210), 100 ell, halofit, wCDM box seed 20240607",
                       "cosmologies_per_gpu": B, "global_cosmologies": world * B, "chunk": chunk,
                       "l2": "per-step working set (workspace %.1f GB + output %.1f GB) exceeds L2; no flush needed"
                             % (ws_bytes / 1e9, out.numel() * 8 / 1e9),
                       "parallelism": "cosmology shards, one rank per GPU, no data-path collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(stage_n.values())),
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
