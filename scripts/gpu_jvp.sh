#!/bin/bash
# usage: scripts/gpu_jvp.sh <tag>  -- forward-mode tests, config4 with / without the reverse-sweep K3, launch list of a config4 step
TAG=$1; O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "jvp or fisher or vjp or autograd or likelihood or hessian" > $O/${TAG}_pytest.log 2>&1; echo rc=$? >> $O/${TAG}_pytest.log
tail -15 $O/${TAG}_pytest.log
for A in 1 0; do
  JC_JVP_ADJOINT=$A timeout 200 python bench.py --workload config4 --steps 5 --warmup 3 > $O/${TAG}_config4_adj$A.json 2> $O/${TAG}_config4_adj$A.err
  python -c "
import json;d=json.loads(open('$O/${TAG}_config4_adj$A.json').read().strip().splitlines()[-1]);print('adjoint', $A, 'ms', d['ms_per_step'], 'jvp/fwd', d['jvp_over_forward'], 'e2e', d['e2e']['value'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/${TAG}_launches_config4.csv python bench.py --workload config4 --steps 1 --warmup 1 > $O/${TAG}_launches_config4.log 2>&1; echo ncu rc=$?
