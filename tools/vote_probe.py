#!/usr/bin/env python
"""One voting launch of a workload on a strided subset of its reference points: kernel time and work counters.

usage: python tools/vote_probe.py [--workload c3] [--ref-step 50] [--repeat 3]
Used for tuning comparisons (B200PPF_LIB=variants/X.so) and for `ncu --set full` captures of the full-size
table with a launch short enough to replay (-k regex:ppf_vote_kernel -s 1 -c 1).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ppf_pose_estimation_b200 import capi, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--ref-step", type=int, default=50)
    ap.add_argument("--ref-first", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    wl = workloads.load(args.workload)
    ctx = capi.Context(0)
    dm, ds = ctx.upload_cloud(wl.library()[0]), ctx.upload_cloud(wl.scene)
    t = ctx.table_build_from_cloud(dm, wl.angle_step, wl.dist_step)
    step = args.ref_step * wl.ref_rate
    count = (wl.scene.shape[0] - args.ref_first + step - 1) // step
    ms = []
    for _ in range(args.repeat):
        hy = ctx.vote(dm, t, ds, args.ref_first, step, count)
        ms.append(ctx.timings()["vote_ms"])
    st = ctx.vote_stats()
    best = min(ms)
    print(json.dumps({"tag": args.tag, "lib": os.environ.get("B200PPF_LIB", "default"), "workload": wl.name, "refs": int(count),
                      "slices": int(t.info.n_slices), "n_alpha": int(t.info.n_alpha), "vote_ms": ms, "votes": st["votes"],
                      "pairs_in_radius": st["pairs_in_radius"], "gvotes_per_s": st["votes"] / (best * 1e-3) / 1e9,
                      "checksum": int(hy["votes"].astype("uint64").sum())}))


if __name__ == "__main__":
    main()
