"""profiles/ncu_current.json from `ncu --set full` captures of the pipeline kernels (one launch each, bench tracer set,
scripts/ncu_target.py: 592 cosmologies).  bench.py reads the file for `roofline.traffic` and the executed-FP64-work figures
(`pipe_frac*`) and refuses it when the kernel sources changed after the capture (sha256 over csrc/*.cu, *.cuh).

    python scripts/make_ncu_current.py <capture tag> <n_cosmo> setup=a.ncu-rep lens=b.ncu-rep finish=... power=... contract=...
    python scripts/make_ncu_current.py <capture tag> <n_cosmo> pass.ncu-rep        (one report with one pass of all five kernels)

Per kernel: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), FP64-datapath work in flop =
2 x (DFMA + DMUL + DADD thread instructions) + sm__ops_path_tensor_src_fp64 (DMMA), duration, pipe / issue utilisation as ncu
saw them (cold cache, serialised -- the bench recomputes utilisation from its own CUDA-event times)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,   # -> bytes
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}                                 # -> microseconds


def raw(rep):
    """Rows of the report's raw page as dicts; byte and time columns are converted to bytes / microseconds (ncu picks one
    unit per column: Kbyte for a small kernel, Mbyte for a large one)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if u in UNIT_SCALE and (h.startswith("dram__bytes") or h == "gpu__time_duration.sum"):
                try:
                    v = repr(float(v.replace(",", "")) * UNIT_SCALE[u])
                except ValueError:
                    pass
            d[h] = v
        res.append(d)
    return res


def num(d, k, default=0.0):
    try:
        return float(d.get(k, default) or default)
    except ValueError:
        return default


def main():
    from bench import csrc_hash
    tag, n_cosmo = sys.argv[1], int(sys.argv[2])
    kernels = {}
    by_stage = []
    names = {"setup": "jc_setup_kernel", "lens": "jc_lens_kernel", "finish": "jc_tracer_finish_kernel", "power": "jc_power",
             "contract": "jc_contract"}
    for arg in sys.argv[3:]:
        if "=" in arg:
            stage, rep = arg.split("=", 1)
            by_stage.append((stage, raw(rep)))
        else:  # one report holding one pass of the pipeline: launches are assigned to their stage by kernel name
            allk = raw(arg)
            for stage, sub in names.items():
                sel = [d for d in allk if sub in d.get("Kernel Name", "")]
                if sel:
                    by_stage.append((stage, sel))
    for stage, launches in by_stage:
        tot = dict(dram=0.0, flops=0.0, us=0.0, inst=0.0)
        knames = []
        for d in launches:  # a stage may be several launches (lens kernel: one per 10 sources)
            cyc = num(d, "sm__cycles_elapsed.avg")
            thr = sum(num(d, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) for op in ("dfma", "dmul", "dadd")) * cyc
            tot["flops"] += 2.0 * thr + num(d, "sm__ops_path_tensor_src_fp64.sum")
            tot["dram"] += num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
            tot["us"] += num(d, "gpu__time_duration.sum")
            tot["inst"] += num(d, "smsp__inst_executed.sum")
            knames.append(d.get("Kernel Name", "?"))
        d0 = launches[0]
        kernels[stage] = {
            "kernel": knames[0], "launches": len(launches),
            "dram_bytes_per_cosmology": tot["dram"] / n_cosmo,
            "fp64_flops_per_cosmology": tot["flops"] / n_cosmo,
            "warp_instructions_per_cosmology": tot["inst"] / n_cosmo,
            "ncu_duration_us": tot["us"],
            "ncu_fp64_pipe_pct": num(d0, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "ncu_dmma_pipe_pct": num(d0, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
            "ncu_issue_active_pct": num(d0, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": num(d0, "launch__registers_per_thread"),
            "ncu_sm_mhz": 1e3 * num(d0, "sm__cycles_elapsed.avg.per_second"),
        }
    rec = {"capture": tag, "n_cosmologies": n_cosmo, "csrc_sha256": csrc_hash(),
           "command": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> python scripts/ncu_target.py",
           "kernels": kernels}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "ncu_current.json"), "w"), indent=1)
    for k, v in kernels.items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})


if __name__ == "__main__":
    main()
