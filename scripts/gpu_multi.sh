#!/bin/bash
# final multi-GPU evidence: 2-GPU tests, then the default bench (as the driver launches it) at N = 8, 4, 2
TAG=$1
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "nccl or two_devices or peer_gather or sharded" > $O/${TAG}_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_multigpu.log
tail -3 $O/${TAG}_pytest_multigpu.log
run() { name=$1; n=$2; shift; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n --steps 10 --warmup 3 "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  echo "$name rc=$?"; }
run n8 8
run n8_push1184 8 --no-e2e --push-rows 1184
run n4 4 --no-e2e
run n2 2 --no-e2e
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_n*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get("gather") or {}
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in g.items() if k in ("mode","lockstep","sub_chunk","push_rows","ms_per_step_compute_only","exposed_ms","ratio_vs_compute_only","exchange_alone_ms","nvlink_in_gbs_alone","bitwise_equal_to_local")}, "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unparsed", e, open(f.replace(".json",".err")).read()[-500:])
P
