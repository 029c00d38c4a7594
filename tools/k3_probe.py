"""C3 vote kernel on one GPU: static grid and shared-queue mode, three runs each (vote_ms from CUDA events)."""
import sys, time, os; sys.path.insert(0, ".")
import ctypes as C
from yolo_ppf_pose_estimation_b200 import capi, workloads
wl = workloads.load(sys.argv[1] if len(sys.argv) > 1 else "c3")
L = capi.lib()
ctx = capi.Context(0); dm, ds = ctx.upload_cloud(wl.model), ctx.upload_cloud(wl.scene)
t = ctx.table_build_from_cloud(dm, wl.angle_step, wl.dist_step)
for _ in range(3):
    f, p, v = ctx.register(dm, t, ds, ref_rate=wl.ref_rate, pos_thr=wl.pos_thr, rot_thr=wl.rot_thr)
    print("static grid: vote_ms", round(ctx.timings()["vote_ms"], 1), v[:2], flush=True)
del t, dm, ds, ctx
m = capi.Multi([0]); m.train(wl.model, wl.angle_step, wl.dist_step); m.scene(wl.scene)
for _ in range(3):
    f, p, v = m.register(ref_rate=wl.ref_rate, pos_thr=wl.pos_thr, rot_thr=wl.rot_thr)
    tm = capi.Timings(); L.b200ppf_get_timings(L.b200ppf_multi_context(m._h, 0), C.byref(tm))
    print("queue mode: vote_ms", round(tm.vote_ms, 1), v[:2], flush=True)
