import sys; sys.path.insert(0, ".")
import numpy as np
from yolo_ppf_pose_estimation_b200 import capi
from oracle import binding as ob
g = "tests/golden/"
import glob
bottle = np.load(g + "bottle_5mm.npz")["cloud"].astype(np.float32)
crop = np.load(g + "scene_crop_1cm.npz")["cloud"].astype(np.float32)
ctx = capi.Context(0)
A = np.float32(12.0) / np.float32(180.0) * np.float32(np.pi); D = np.float32(0.01)
feats = ob.ppf_estimation(bottle)
hm = ob.HashMap(A, D).set_input_feature_cloud(feats)
t = ctx.table_build(ctx.features_upload(feats), A, D)
ds = ctx.upload_cloud(crop)
for s_r in (0, 433, 933):
    inr, d, a = ctx.vote_debug_pairs(t, ds, s_r)
    acc = ctx.vote_debug_accumulator(t, ds, s_r)
    ref, votes = hm.vote_accumulate_from_pairs(bottle.shape[0], d[inr > 0], a[inr > 0])
    diff = acc.astype(np.int64) - ref.astype(np.int64)
    nz = np.argwhere(diff != 0)
    print("ref", s_r, "votes", votes, int(acc.sum()), "cells differing", len(nz), "abs diff sum", int(np.abs(diff).sum()))
    for r, c in nz[:12]:
        print("   row", r, "bin", c, "got", acc[r, c], "want", ref[r, c])
