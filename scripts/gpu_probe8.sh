#!/bin/bash
TAG=$1
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/gather_probe.py > $O/${TAG}_probe.log 2> $O/${TAG}_probe.err
echo rc=$?; cat $O/${TAG}_probe.log; tail -5 $O/${TAG}_probe.err
JC_GATHER_STREAMS=4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > $O/${TAG}_n8_streams4.json 2> $O/${TAG}_n8_streams4.err
JC_GATHER_STREAMS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > $O/${TAG}_n8_streams1.json 2> $O/${TAG}_n8_streams1.err
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_n8*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get("gather") or {}
        print(f, "ms %.3f" % d["ms_per_step"], g.get("exposed_ms"))
    except Exception as e:
        print(f, "unparsed", e)
P
