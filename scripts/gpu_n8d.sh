#!/bin/bash
TAG=$1
O=gpurun_out
run() { name=$1; n=$2; shift; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 10 --warmup 3 --no-e2e "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  echo "$name rc=$?"; }
export JC_GATHER_STREAMS=2
run n8_lock_s2 8 --gather-mode peer_ce
export JC_GATHER_STREAMS=1
run n8_lock_s1 8 --gather-mode peer_ce
export JC_GATHER_STREAMS=2
JC_GATHER_LOCKSTEP=0 run n8_free_s2 8 --gather-mode peer_ce
run n8_sms16 8 --gather-mode peer --push-sms 16
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_n*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get("gather") or {}
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in g.items() if k in ("mode","push_sms","sub_chunk","push_rows","ms_per_step_compute_only","exposed_ms","ratio_vs_compute_only","exchange_alone_ms","nvlink_in_gbs_alone","bitwise_equal_to_local")})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace(".json",".err")).read()[-500:])
P
