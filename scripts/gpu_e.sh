#!/bin/bash
TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err; }
run base JC_X=0
run tab1 JC_POWER_TAB_NPT=-1
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 > $O/${TAG}_config3.json 2> $O/${TAG}_config3.err; echo "config3 rc=$?"; tail -3 $O/${TAG}_config3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jc_power_tab2 -s 1 -c 1 -f -o $O/${TAG}_jc_power_tab2 python scripts/ncu_target.py > $O/${TAG}_ncu_tab2.log 2>&1
timeout 600 python scripts/parity_sweep.py 256 > $O/${TAG}_parity_sweep.log 2>&1; tail -2 $O/${TAG}_parity_sweep.log
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:round(v,3) for k,v in r.items() if k.startswith("ms_")})
    except Exception as e:
        print(f, "unparsed", e)
for f in ("gpurun_out/%s_config3.json"%tag,):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["launch_ms"], d["explicit_covariance_form"])
    except Exception as e: print(f,"unparsed",e)
P
