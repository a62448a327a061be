#!/bin/bash
# usage: scripts/gpu_prof.sh <tag>  -- GPU tests, N=1 bench, launch list, ncu --set full of the power and contraction kernels
TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"; tail -2 $O/${TAG}_bench_n1.err
JC_CONTRACT_EPS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_eps0.json 2> $O/${TAG}_bench_eps0.err
for K in jc_power_tab jc_contract_tma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o $O/${TAG}_$K python scripts/ncu_target.py > $O/${TAG}_ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --peak-tflops 36.4 > $O/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/*_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value",0)), {k:round(v,3) for k,v in r.items() if k.startswith("ms_")})
    except Exception as e:
        print(f, "unparsed", e)
P
