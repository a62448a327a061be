#!/bin/bash
TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
for NPT in 4 16; do
JC_POWER_TAB_NPT=$NPT timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_npt$NPT.json 2> $O/${TAG}_bench_npt$NPT.err
done
JC_POWER_EXACT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_exact.json 2> $O/${TAG}_bench_exact.err
timeout 900 python scripts/parity_sweep.py 1024 > $O/${TAG}_parity_sweep.log 2>&1; tail -5 $O/${TAG}_parity_sweep.log
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
