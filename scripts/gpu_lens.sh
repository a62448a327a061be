#!/bin/bash
# usage: scripts/gpu_lens.sh <tag>  -- the DMMA lens kernel: its test, the whole GPU suite with it switched on, bench A/B, ncu
TAG=$1; O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "lens_mma" > $O/${TAG}_pytest_lens.log 2>&1; echo "lens test rc=$?"; tail -5 $O/${TAG}_pytest_lens.log
JC_LENS_MMA=1 timeout 600 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_all_mma.log 2>&1; echo "suite with lens_mma rc=$?"; tail -4 $O/${TAG}_pytest_all_mma.log
for A in 0 1; do
  JC_LENS_MMA=$A timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_mma$A.json 2> $O/${TAG}_bench_mma$A.err
  python -c "
import json;d=json.loads(open('$O/${TAG}_bench_mma$A.json').read().strip().splitlines()[-1]);r=d['roofline'];print('lens_mma', $A, 'ms', round(d['ms_per_step'],3), {k:round(v,3) for k,v in r.items() if k.startswith('ms_')})"
done
JC_LENS_MMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:jc_lens_mma -s 1 -c 1 -f -o $O/${TAG}_lensmma python scripts/ncu_target.py > $O/${TAG}_ncu.log 2>&1; echo ncu rc=$?
