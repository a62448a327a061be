#!/bin/bash
# Tuning sweep over the JC_POWER_CFG / JC_CONTRACT_CFG kernel variants (stage ms per step of 8192 cosmologies).
for p in 0 1 2 3 4 5 6; do
  JC_POWER_CFG=$p python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms_per_step']
print('power_cfg $p  power %.2f ms  step %.2f ms' % (s['power'], d['ms_per_step']))"
done
for c in 0; do
  JC_CONTRACT_CFG=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stage_ms_per_step']
print('contract_cfg $c  contract %.2f ms  step %.2f ms' % (s['contract'], d['ms_per_step']))"
done
