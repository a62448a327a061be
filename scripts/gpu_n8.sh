#!/bin/bash
# 8-GPU box: the gather-inclusive bench at N = 8, 4, 2 (peer pushes), the NCCL send/recv pipeline at N = 8, two pipeline shapes
TAG=$1
O=gpurun_out
nvidia-smi topo -m > $O/${TAG}_topo.log 2>&1
run() { name=$1; n=$2; shift; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 10 --warmup 3 --no-e2e "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  echo "$name rc=$?"; }
run n8_peer 8
run n8_nccl 8 --gather-mode nccl
run n8_peer_push296 8 --push-rows 296
run n8_peer_sub2368 8 --sub-chunk 2368
run n4_peer 4
run n2_peer 2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 10 --warmup 3 > $O/${TAG}_n8_full.json 2> $O/${TAG}_n8_full.err; echo "n8 full rc=$?"
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_n*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get("gather") or {}
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in g.items() if k in ("mode","sub_chunk","push_rows","ms_per_step_compute_only","exposed_ms","ratio_vs_compute_only","exchange_alone_ms","nvlink_in_gbs_alone","bitwise_equal_to_local")}, "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unparsed", e, open(f.replace(".json",".err")).read()[-600:])
P
