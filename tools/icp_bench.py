#!/usr/bin/env python
"""K6 timing: the reference's ICP(100, 0.005, 2.5, 8).registerModelToScene on five start poses — B200 vs the CPU
restatement (one thread per pose sequentially, as a plain loop).  Prints one JSON line.

usage: python tools/icp_bench.py [--no-cpu] [--repeat 5]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()
    from yolo_ppf_pose_estimation_b200 import capi
    from test_gpu_parity import _icp_case, axis_pose_error
    model, scene, G, starts = _icp_case()
    c = capi.Context(0)
    dm, ds = c.upload_cloud(model), c.upload_cloud(scene)
    c.icp_refine(dm, ds, starts)  # warm-up
    ms = []
    for _ in range(args.repeat):
        P, res, it = c.icp_refine(dm, ds, starts)
        ms.append(c.timings()["icp_ms"])
    rec = {"stage": "K6 icp_refine", "n_model": int(model.shape[0]), "n_scene": int(scene.shape[0]), "poses": len(starts),
           "params": [100, 0.005, 2.5, 8], "iterations": it, "kernel_ms": float(np.median(ms)),
           "ms_per_iteration": float(np.median(ms)) / max(1, it) * len(starts),
           "errors_vs_ground_truth_m_deg": [axis_pose_error(P[k], G) for k in range(len(starts))]}
    if not args.no_cpu:
        from oracle import binding as ob
        t0 = time.perf_counter()
        R, rres, rit = ob.icp_refine(model, scene, starts)
        rec["cpu_ms_one_thread"] = 1e3 * (time.perf_counter() - t0)
        rec["cpu_iterations"] = rit
        rec["max_translation_diff_m"] = float(max(np.linalg.norm(P[k][:3, 3] - R[k][:3, 3]) for k in range(len(starts))))
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
