#!/bin/bash
# usage (gpurun --gpus 2): scripts/gpu_n2_final.sh <tag>  -- the multi-GPU tests and the 2-GPU bench line (gather inside the timed step)
TAG=$1; O=gpurun_out
nvidia-smi --query-gpu=name --format=csv > $O/${TAG}_smi.log 2>&1
timeout 400 python -m pytest tests -m gpu -q -k "nccl or two_devices or peer or sharded" > $O/${TAG}_pytest_multigpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_multigpu.log
tail -4 $O/${TAG}_pytest_multigpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; tail -2 $O/${TAG}_bench_n2.err
python - <<P
import json
d=json.loads(open("$O/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), json.dumps(d.get("gather"))[:700], "e2e", (d.get("e2e") or {}).get("value"))
P
