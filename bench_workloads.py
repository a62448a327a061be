"""bench.py --workload config3 | config4: the two BASELINE configurations that consume the spectra on the device.

config3  3x2pt 10+10 bins, 100 ell, halofit: per cosmology the reference's
             mu, cov = gaussian_cl_covariance_and_mean(cosmo, ell, probes, sparse=True); lnL = gaussian_log_likelihood(data, mu, cov)
         as one device pipeline -- K1..K4 -> jc_gaussian_cl_loglike_f64 (the covariance is never formed: T x T identity,
         csrc/jc_cl_loglike.cu) -- B doubles leave the device.  `value` / `e2e` count C_ell evaluations (cosmology x ell x
         pair) per second like the headline metric, so the numbers compare directly with config 5; `likelihoods_per_s`
         is the same rate per cosmology.  For reference the explicit two-kernel form (jc_gaussian_cov_f64 ->
         jc_gaussian_loglike_f64: 35 MB of covariance written and read back per cosmology) is timed on a sub-batch, with the
         covariance kernel's achieved HBM write rate against MEASURED_PEAKS.json.
config4  3x2pt 5+5 bins, 100 ell, halofit, forward-mode derivatives with respect to the 7 wCDM parameters for a batch of
         cosmologies: (cl [B,55,100], dcl [B,7,55,100]) per step; unit dC_ell/s = derivative entries per second
         (cosmology x parameter x ell x pair).
"""
import json
import os
import time

import numpy as np

N_ELL = 100
METRIC = "C_ell evals/sec (cosmo x ell x pair)"


def run(args, rank, world, local):
    import torch
    import torch.distributed as dist

    import jax_cosmo_b200 as jc
    from bench import ClockSampler, git_head
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    steps, warmup = args.steps, max(args.warmup, 3)
    box = sc.config5_cosmologies(65536)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    if args.workload == "config3":
        n_src = n_lens = 10
        T = n_src + n_lens
        P = T * (T + 1) // 2
        B = args.cosmologies_per_gpu or 8192
        scn = sc.scenario("cfg3", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(n_src, 1.0), sc.lenses(n_lens, 1.0)])
        probes = sc.build_probes(scn, jc)
        plan = _native.get_plan(probes, scn["ell"], None, None, device=local)
        rows = np.ascontiguousarray(box[(rank * B) % 65536:][:B] if (rank * B) % 65536 + B <= 65536 else box[:B])
        cos = torch.as_tensor(rows, device=dev)
        cl = torch.empty((B, P, N_ELL), dtype=torch.float64, device=dev)
        ws = plan.workspace(B)
        # the data vector: the fiducial (Planck15) spectra with 1 % multiplicative scatter, fixed seed
        fid = plan.angular_cl_device(torch.as_tensor(sc.cosmo_row(sc.PLANCK15)[None], device=dev))[0]
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        data = (fid * (1.0 + 0.01 * torch.randn(fid.shape, generator=g, dtype=torch.float64, device=dev))).reshape(-1).contiguous()

        def step():
            plan.angular_cl_device(cos, out=cl, workspace=ws)
            return plan.gaussian_cl_loglike_device(cl, data, 0.25)

        for _ in range(warmup):
            lnl = step()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        e0.record()
        for _ in range(steps):
            lnl = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        # the likelihood stage by itself
        barrier()
        e1.record()
        for _ in range(steps):
            plan.gaussian_cl_loglike_device(cl, data, 0.25)
        e2.record()
        barrier()
        ms_like = e1.elapsed_time(e2) / steps
        clocks = sampler.stop()
        evals = world * B * P * N_ELL
        value = evals * steps / (ms * 1e-3)
        # e2e: host rows in (pinned), B log-likelihoods out
        rows_pin = torch.as_tensor(rows).pin_memory()
        out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
        cos2 = torch.empty_like(cos)

        def step_e2e():
            cos2.copy_(rows_pin, non_blocking=True)
            plan.angular_cl_device(cos2, out=cl, workspace=ws)
            out_pin.copy_(plan.gaussian_cl_loglike_device(cl, data, 0.25), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": evals * steps / dt, "unit": "C_ell/s", "h2d_bytes_per_step": int(rows.nbytes), "d2h_bytes_per_step": 8 * B,
               "max_abs_diff_vs_device_path": float((out_pin.to(dev) - lnl).abs().max().item()), "rows_compared": B}
        # the explicit two-kernel form on a sub-batch (35.3 MB of covariance per cosmology)
        nb = 256
        cov_bytes = nb * P * P * N_ELL * 8
        noise_dev = torch.as_tensor(plan.noise(), device=dev)
        cov = plan.gaussian_cov_device(cl[:nb], 0.25)
        mu = cl[:nb].reshape(nb, -1)
        ref = _native.gaussian_loglike_device(data, mu, cov)
        agree = float(((ref - lnl[:nb]).abs() / ref.abs()).max().item())
        torch.cuda.synchronize()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record()
        for _ in range(3):
            check = _native.load_library().jc_gaussian_cov_f64(plan._h, cl[:nb].data_ptr(), noise_dev.data_ptr(), nb, 0.25, cov.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream)
        c1.record()
        for _ in range(3):
            _native.gaussian_loglike_device(data, mu, cov)
        c2.record()
        torch.cuda.synchronize()
        ms_cov, ms_chol = c0.elapsed_time(c1) / 3, c1.elapsed_time(c2) / 3
        if rank == 0:
            like_bytes = B * P * N_ELL * 8.0
            line = {"metric": METRIC, "value": value, "unit": "C_ell/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                    "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic",
                    "config": {"workload": "config3: 3x2pt 10+10 Smail bins (T=20, P=210), 100 ell, halofit; spectra -> Gaussian "
                                           "log-likelihood under the covariance of the model spectra, on the device",
                               "cosmologies_per_gpu": B, "l2": "per-step working set exceeds L2; no flush needed"},
                    "likelihoods_per_s": world * B * steps / (ms * 1e-3),
                    "clocks": clocks, "e2e": e2e, "gpu_launches": (10 + 2) * steps,
                    "roofline": {"bound": "hbm", "kernel": "jc_cl_loglike_kernel", "achieved": like_bytes / (ms_like * 1e-3) / 1e9,
                                 "peak": hbm_peak, "unit": "GB/s", "frac": like_bytes / (ms_like * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                 "peak_source": hbm_src, "launch_ms": ms_like,
                                 "note": "the fused likelihood reads the [B,210,100] spectra once (algorithmic bytes 168 KB per "
                                         "cosmology); it is %.1f %% of the step -- the step itself is the FP64-bound K1..K4 pipeline "
                                         "of config 5" % (100.0 * ms_like / (ms / steps))},
                    "explicit_covariance_form": {
                        "cosmologies": nb, "cov_kernel_ms": ms_cov, "cov_write_gbs": cov_bytes / (ms_cov * 1e-3) / 1e9,
                        "cov_write_frac_of_hbm_peak": cov_bytes / (ms_cov * 1e-3) / 1e9 / hbm_peak,
                        "cholesky_kernel_ms": ms_chol, "ms_per_cosmology": (ms_cov + ms_chol) / nb,
                        "fused_ms_per_cosmology": ms_like / B, "max_rel_diff_fused_vs_explicit": agree,
                        "note": "jc_gaussian_cov_f64 -> jc_gaussian_loglike_f64: 35.3 MB of sparse covariance written and read "
                                "back per cosmology, P x P Cholesky per ell"},
                    "cpu_baseline": None, "git_head": git_head()}
            print(json.dumps(line), flush=True)
    else:  # config4
        n_src = n_lens = 5
        T = n_src + n_lens
        P = T * (T + 1) // 2
        K = 7
        B = args.cosmologies_per_gpu or 1024
        scn = sc.scenario("cfg4", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(n_src, 2.0), sc.lenses(n_lens, 2.0)])
        probes = sc.build_probes(scn, jc)
        plan = _native.get_plan(probes, scn["ell"], None, None, device=local)
        rows = np.ascontiguousarray(box[(rank * B) % 65536:][:B] if (rank * B) % 65536 + B <= 65536 else box[:B])
        cos = torch.as_tensor(rows, device=dev)
        tang = torch.zeros((K, 8), dtype=torch.float64, device=dev)
        # the 7 wCDM directions, those the tracer kernels depend on first (Omega_c, Omega_b, w0, wa | h, n_s, sigma8): the second
        # tangent group then has dR = 0 and its K2 pass is skipped (_native.direction_order does this for host-built tangents)
        tang[torch.arange(K), torch.tensor([0, 1, 6, 7, 2, 3, 4])] = 1.0
        for _ in range(warmup):
            cl, dcl = plan.angular_cl_jvp_device(cos, tang)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            cl, dcl = plan.angular_cl_jvp_device(cos, tang)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop()
        # forward pass of the same batch for the cost ratio
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan.angular_cl_device(cos)
        f0.record()
        for _ in range(steps):
            plan.angular_cl_device(cos)
        f1.record()
        torch.cuda.synchronize()
        ms_fwd = f0.elapsed_time(f1) / steps
        # e2e: host rows in, Jacobian out (pinned)
        out_pin = torch.empty((B, K, P, N_ELL), dtype=torch.float64).pin_memory()
        rows_pin = torch.as_tensor(rows).pin_memory()

        def step_e2e():
            c = rows_pin.to(dev, non_blocking=True)
            _, d = plan.angular_cl_jvp_device(c, tang)
            out_pin.copy_(d, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        dt = max_over_ranks(time.perf_counter() - t0)
        derivs = world * B * K * P * N_ELL
        # kernels per timed step (the throughput path of jc_angular_cl_jvp_f64 in chunks of <= 1024 cosmologies; 5 sources = one
        # lens launch per pass) + the forward pass timed beside it: the claim `gpu_launches` makes
        group, adjoint = int(_native.get_option("jvp_group")), int(_native.get_option("jvp_adjoint"))
        n_chunks = -(-B // 1024)
        if B * K <= 512:
            launches = 4 + 2
        elif adjoint and group == 4 and 3 <= K <= 8:
            launches = n_chunks * (3 * -(-K // 4) + 1 + 1 + K)
        else:
            launches = n_chunks * (4 * -(-K // group) + 1 + K)
        launches *= steps  # the forward pass timed beside it is outside the timed region
        if rank == 0:
            line = {"metric": "dC_ell/d theta evals/sec (cosmo x parameter x ell x pair)", "value": derivs * steps / (ms * 1e-3),
                    "unit": "dC_ell/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": "config4: 3x2pt 5+5 Smail bins (T=10, P=55), 100 ell, halofit, forward-mode d/d(Omega_c, "
                                           "Omega_b, h, n_s, sigma8, w0, wa) for a batch of cosmologies",
                               "direction_order": "Omega_c, Omega_b, w0, wa, h, n_s, sigma8",
                               "cosmologies_per_gpu": B, "tangents": K, "l2": "per-step working set exceeds L2; no flush needed"},
                    "clocks": clocks,
                    "e2e": {"value": derivs * steps / dt, "unit": "dC_ell/s", "h2d_bytes_per_step": int(rows.nbytes),
                            "d2h_bytes_per_step": int(out_pin.numel() * 8),
                            "max_abs_diff_vs_device_path": float((out_pin[:64].to(dev) - dcl[:64]).abs().max().item())},
                    "forward_pass_ms": ms_fwd, "jvp_over_forward": (ms / steps) / ms_fwd,
                    "jvp_group": int(_native.get_option("jvp_group")), "jvp_adjoint": int(_native.get_option("jvp_adjoint")),
                    "gpu_launches": launches,
                    "roofline": None,
                    "roofline_note": "FP64-pipe bound like config 5; per-kernel times of a step: profiles/r02_final_launches_config4.csv "
                                     "(reverse-sweep K3 42 %, tangent contractions 29 %, K1 17 %, lens 9 %), ncu of the K3 kernel: "
                                     "profiles/r02x_ncu_power_adj_metrics.csv (36 % of the FP64 pipe)",
                    "cpu_baseline": None, "git_head": git_head()}
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
