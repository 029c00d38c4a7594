#!/bin/bash
# ncu --set full capture of the voting kernel on one workload.  usage: bash tools/gpu_prof.sh TAG [workload]
TAG=${1:-p}
WL=${2:-c2}
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ppf_vote_kernel -s 3 -c 1 -f -o $O/prof_vote_${WL}_$TAG \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu > $O/ncu_full_${WL}_$TAG.log 2>&1
tail -3 $O/ncu_full_${WL}_$TAG.log | cut -c1-300
