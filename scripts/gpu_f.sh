#!/bin/bash
TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
timeout 300 python bench.py --workload config3 --steps 5 --warmup 3 > $O/${TAG}_config3.json 2> $O/${TAG}_config3.err; echo "config3 rc=$?"; tail -3 $O/${TAG}_config3.err
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in ("gpurun_out/%s_config3.json"%tag,):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["launch_ms"], d["explicit_covariance_form"])
    except Exception as e: print(f,"unparsed",e)
P
