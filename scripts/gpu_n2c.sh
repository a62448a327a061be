#!/bin/bash
TAG=$1
O=gpurun_out
JC_GATHER_LOCKSTEP=2 timeout 600 python -m pytest tests -m gpu -x -q -k "nccl or two_devices or peer_gather or sharded" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
JC_GATHER_LOCKSTEP=2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --gather-mode peer_ce > $O/${TAG}_n2_lock.json 2> $O/${TAG}_n2_lock.err; echo "lock rc=$?"; tail -3 $O/${TAG}_n2_lock.err
python - $TAG <<'P'
import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob("gpurun_out/%s_n*.json"%tag)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get("gather") or {}
        print(f, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in g.items() if k in ("mode","push_sms","sub_chunk","push_rows","ms_per_step_compute_only","exposed_ms","ratio_vs_compute_only","exchange_alone_ms","nvlink_in_gbs_alone","bitwise_equal_to_local","pusher_aborted")})
    except Exception as e:
        print(f, "unparsed", e)
P
