#!/bin/bash
# usage: scripts/gpu_call.sh <tag> <ngpus>   -- GPU tests, N=1 bench, N-GPU bench with the gather (peer / nccl)
TAG=$1; N=${2:-1}
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/${TAG}_smi.log 2>&1
nvidia-smi topo -m >> $O/${TAG}_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
if [ "$N" -gt 1 ]; then
  for MODE in peer nccl; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --gather-mode $MODE > $O/${TAG}_bench_n${N}_${MODE}.json 2> $O/${TAG}_bench_n${N}_${MODE}.err
    echo "bench n$N $MODE rc=$?"; tail -3 $O/${TAG}_bench_n${N}_${MODE}.err
  done
fi
python - <<'P'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/*_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value",0)), json.dumps(d.get("gather"))[:600])
    except Exception as e:
        print(f, "unparsed", e)
P
