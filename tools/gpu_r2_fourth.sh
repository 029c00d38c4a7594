#!/bin/bash
O=gpurun_out; mkdir -p $O; TAG=${1:-v9}
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 | tee $O/r2_pytest_gpu_$TAG.log
# K4 launch list on C3
cat > /tmp/k4probe.py <<PY
import sys; sys.path.insert(0,".")
from yolo_ppf_pose_estimation_b200 import capi, workloads
wl=workloads.load("c3"); ctx=capi.Context(0)
dm,ds=ctx.upload_cloud(wl.model),ctx.upload_cloud(wl.scene)
t=ctx.table_build_from_cloud(dm,wl.angle_step,wl.dist_step)
hy=ctx.vote(dm,t,ds,0,1)
for _ in range(2):
    p,v=ctx.cluster(hy,wl.pos_thr,wl.rot_thr); print("c3 cluster_ms",ctx.timings()["cluster_ms"], v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cluster_\|keep_\|leader_ -c 400 --csv --log-file $O/r2_launches_k4_c3_$TAG.csv python /tmp/k4probe.py > $O/r2_ncu_k4_$TAG.log 2>&1; tail -3 $O/r2_ncu_k4_$TAG.log
python - <<PY
import csv,collections
rows=list(csv.reader(open("$O/r2_launches_k4_c3_$TAG.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
    if len(r)>vi: n=r[ki].split('(')[0][-36:]; agg[n]+=float(r[vi].replace(',','')); cnt[n]+=1
for n,t in agg.most_common(): print(f'{n:40s} {cnt[n]:4d} {t/1e6:9.3f} ms (two cluster calls)')
PY
