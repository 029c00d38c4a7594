#!/bin/bash
O=gpurun_out; mkdir -p $O; TAG=${1:-v10}
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 | tee $O/r2_pytest_gpu_$TAG.log
python /dev/stdin <<PY
import sys; sys.path.insert(0,".")
from yolo_ppf_pose_estimation_b200 import capi, workloads
wl=workloads.load("c3"); ctx=capi.Context(0)
dm,ds=ctx.upload_cloud(wl.model),ctx.upload_cloud(wl.scene)
for _ in range(2):
    t=ctx.table_build_from_cloud(dm,wl.angle_step,wl.dist_step); print("build", {k:round(v,2) for k,v in ctx.timings().items() if k in ("keys_ms","sort_ms","csr_ms")}, t.info.n_merged, t.info.n_entries)
hy=ctx.vote(dm,t,ds,0,1)
for _ in range(3):
    p,v=ctx.cluster(hy,wl.pos_thr,wl.rot_thr); print("c3 cluster_ms",ctx.timings()["cluster_ms"], v)
PY
timeout 900 python tools/table_build_sweep.py --sizes 5000,10000,20000,35000,50000 --features --cpu-max 10000 > $O/r2_table_build_sweep_$TAG.jsonl 2> $O/r2_table_build_sweep_$TAG.err; cat $O/r2_table_build_sweep_$TAG.jsonl | cut -c1-700; tail -3 $O/r2_table_build_sweep_$TAG.err
