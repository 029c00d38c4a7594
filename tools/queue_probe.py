import sys, time; sys.path.insert(0,".")
from yolo_ppf_pose_estimation_b200 import capi, workloads
import ctypes as C
wl=workloads.load("c3")
m=capi.Multi([0]); m.train(wl.model, wl.angle_step, wl.dist_step); m.scene(wl.scene)
L=capi.lib()
for _ in range(3):
    t0=time.perf_counter(); f,p,v=m.register(ref_rate=1, pos_thr=wl.pos_thr, rot_thr=wl.rot_thr); dt=time.perf_counter()-t0
    tm=capi.Timings(); L.b200ppf_get_timings(L.b200ppf_multi_context(m._h,0), C.byref(tm))
    print("queue mode world=1: wall", round(dt*1e3,1), "vote_ms", round(tm.vote_ms,1), "cluster_ms", round(tm.cluster_ms,1), v)
ctx=capi.Context(0); dm,ds=ctx.upload_cloud(wl.model),ctx.upload_cloud(wl.scene); t=ctx.table_build_from_cloud(dm,wl.angle_step,wl.dist_step)
for _ in range(3):
    t0=time.perf_counter(); f,p,v=ctx.register(dm,t,ds,ref_rate=1,pos_thr=wl.pos_thr,rot_thr=wl.rot_thr); dt=time.perf_counter()-t0
    print("static grid: wall", round(dt*1e3,1), "vote_ms", round(ctx.timings()["vote_ms"],1), "cluster_ms", round(ctx.timings()["cluster_ms"],1), v)
