#!/bin/bash
TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err; }
run base JC_X=0
run noskip JC_CONTRACT_EPS=-1
run pair1 JC_CONTRACT_PAIRING=1
run pair2 JC_CONTRACT_PAIRING=2
run ring16 JC_CONTRACT_CFG=2
run ring16pair1 JC_CONTRACT_CFG=2 JC_CONTRACT_PAIRING=1
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 > $O/${TAG}_config3.json 2> $O/${TAG}_config3.err; echo "config3 rc=$?"; tail -3 $O/${TAG}_config3.err
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 > $O/${TAG}_config4.json 2> $O/${TAG}_config4.err; echo "config4 rc=$?"; tail -3 $O/${TAG}_config4.err
python - <<'P'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/r02d_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:round(v,3) for k,v in r.items() if k.startswith("ms_")})
    except Exception as e:
        print(f, "unparsed", e)
for f in ("gpurun_out/r02d_config3.json","gpurun_out/r02d_config4.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, json.dumps(d)[:1800])
    except Exception as e: print(f,"unparsed",e)
P
