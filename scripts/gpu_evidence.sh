#!/bin/bash
# usage: scripts/gpu_evidence.sh <tag> [final]
#   always: GPU tests, N=1 bench line, config4 with tangent groups 4 and 1, config3, one ncu --set full pass over the five
#           pipeline kernels -> ncu_current.json, launch list of a config4 step
#   final:  also the CPU arm in the N=1 line, the bench line that reads the fresh ncu_current.json, the launch list of the bench
#           command and `--impl reference`
TAG=$1; MODE=$2
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
CPUARM=--no-cpu-baseline; [ "$MODE" = final ] && CPUARM=
timeout 300 python bench.py --steps 10 --warmup 3 $CPUARM > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"; tail -2 $O/${TAG}_bench_n1.err
for G in 4 1; do
  JC_JVP_GROUP=$G timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 > $O/${TAG}_config4_g$G.json 2> $O/${TAG}_config4_g$G.err; echo "config4 g$G rc=$?"
done
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 > $O/${TAG}_config3.json 2> $O/${TAG}_config3.err; echo "config3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:jc_(setup|lens|tracer_finish|power_tab2|contract_tma)_kernel' -s 5 -c 5 -f -o $O/${TAG}_pass python scripts/ncu_target.py > $O/${TAG}_ncu_pass.log 2>&1
echo "ncu pass rc=$?"; tail -3 $O/${TAG}_ncu_pass.log
python scripts/make_ncu_current.py $TAG 592 $O/${TAG}_pass.ncu-rep > $O/${TAG}_ncu_current.log 2>&1 && cp profiles/ncu_current.json $O/${TAG}_ncu_current.json
tail -6 $O/${TAG}_ncu_current.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_config4.csv python bench.py --workload config4 --steps 1 --warmup 1 > $O/${TAG}_launches_config4.log 2>&1
echo "launch list config4 rc=$?"
if [ "$MODE" = final ]; then
  timeout 600 python scripts/parity_sweep.py 1024 > $O/${TAG}_parity_sweep.log 2>&1; echo "parity sweep rc=$?"; tail -2 $O/${TAG}_parity_sweep.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1_ncu.json 2> $O/${TAG}_bench_n1_ncu.err; echo "bench (with ncu_current) rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --peak-tflops 36.4 > $O/${TAG}_launches_bench.log 2>&1
  echo "launch list rc=$?"
  timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_reference_arm.json 2> $O/${TAG}_reference_arm.err; echo "reference arm rc=$?"
fi
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_*.json" % tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value",0)),
              {k:round(v,3) for k,v in r.items() if k.startswith("ms_") or (k.startswith("pipe_frac") and isinstance(v,float))},
              "jvp/fwd", d.get("jvp_over_forward"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unparsed", e)
P
