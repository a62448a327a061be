#!/bin/bash
# usage: scripts/gpu_jvp_adjoint.sh <tag>  -- forward-mode tests, config4 with / without the reverse-sweep K3, ncu --set full of that kernel
TAG=$1; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "jvp" > $O/${TAG}_pytest.log 2>&1; echo rc=$? >> $O/${TAG}_pytest.log; tail -6 $O/${TAG}_pytest.log
for A in 1 0; do
  JC_JVP_ADJOINT=$A timeout 200 python bench.py --workload config4 --steps 5 --warmup 3 > $O/${TAG}_config4_adj$A.json 2> $O/${TAG}_config4_adj$A.err
  python -c "
import json;d=json.loads(open('$O/${TAG}_config4_adj$A.json').read().strip().splitlines()[-1]);print('adjoint', $A, 'ms', d['ms_per_step'], 'jvp/fwd', d['jvp_over_forward'], 'e2e', d['e2e']['value'])"
done
export JC_JVP_ADJOINT=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jc_power_adj -s 1 -c 1 -f -o $O/${TAG}_adj python bench.py --workload config4 --steps 1 --warmup 1 --cosmologies-per-gpu 592 > $O/${TAG}_ncu.log 2>&1; echo ncu rc=$?
