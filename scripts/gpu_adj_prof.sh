#!/bin/bash
# usage: scripts/gpu_adj_prof.sh <tag>  -- ncu of the reverse-sweep K3 kernel inside a config4 step
TAG=$1; O=gpurun_out
export JC_JVP_ADJOINT=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_config4_adj.csv python bench.py --workload config4 --steps 1 --warmup 1 --cosmologies-per-gpu 592 > $O/${TAG}_l.log 2>&1; echo launches rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jc_power_adj -s 1 -c 1 -f -o $O/${TAG}_adj python bench.py --workload config4 --steps 1 --warmup 1 --cosmologies-per-gpu 592 > $O/${TAG}_ncu.log 2>&1; echo ncu rc=$?
tail -3 $O/${TAG}_ncu.log
