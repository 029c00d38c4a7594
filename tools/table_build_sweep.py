#!/usr/bin/env python
"""C5 of BASELINE.json: model hash-table build sweep on one B200 (K1+K2 fused, and K1 alone).

usage: python tools/table_build_sweep.py [--sizes 5000,10000,...] [--features]
Prints one JSON line per model size: pairs, stage times (CUDA events inside the library), pairs/s,
and the HBM-roofline fraction of the fused build (algorithmic bytes: 8 B written per pair by the key
kernel + 12 B of CSR entries per pair; the radix passes above that are overhead — SURVEY.md §8d).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yolo_ppf_pose_estimation_b200 import capi, synth, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="5000,10000,20000,35000,50000")
    ap.add_argument("--features", action="store_true", help="also time K1 with the materialised N*N*20 B feature cloud")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--cpu-max", type=int, default=10000,
                    help="largest model for which the CPU side (PPFEstimation + PPFHashMapSearch::setInputFeatureCloud, one thread = PCL "
                         "as shipped) is measured; larger ones are extrapolated in proportion to the pairs and flagged")
    args = ap.parse_args()
    peak = 6549.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    ctx = capi.Context(0)
    cpu_ref = None  # (pairs, estimation seconds, container seconds) of the largest measured CPU build
    for n in [int(x) for x in args.sizes.split(",")]:
        model = synth.synth_model(n, 1)
        dm = ctx.upload_cloud(model)
        best = None
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            table = ctx.table_build_from_cloud(dm, workloads.ANGLE_STEP, workloads.DIST_STEP)
            wall = 1e3 * (time.perf_counter() - t0)
            tim = ctx.timings()
            info = table.info
            rec = {"n_model": n, "pairs": n * (n - 1), "entries": int(info.n_entries), "keys": int(info.n_keys),
                   "key_bits": int(info.key_bits), "slices": int(info.n_slices), "wall_ms": wall,
                   "keys_ms": tim["keys_ms"], "sort_ms": tim["sort_ms"], "csr_ms": tim["csr_ms"]}
            dev_ms = rec["keys_ms"] + rec["sort_ms"] + rec["csr_ms"]
            rec["device_ms"] = dev_ms
            rec["pairs_per_sec"] = rec["pairs"] / (dev_ms * 1e-3)
            rec["algorithmic_gbs"] = 20.0 * rec["pairs"] / (dev_ms * 1e-3) / 1e9
            rec["hbm_frac"] = rec["algorithmic_gbs"] / peak
            table.free()
            if best is None or rec["device_ms"] < best["device_ms"]:
                best = rec
        if args.features and n <= 30000:
            t0 = time.perf_counter()
            F = ctx.features_compute(dm)
            best["k1_features_ms"] = ctx.timings()["features_ms"]
            best["k1_write_gbs"] = 20.0 * n * n / (best["k1_features_ms"] * 1e-3) / 1e9
            best["k1_hbm_frac"] = best["k1_write_gbs"] / peak
            F.free()
        # the CPU side of BASELINE config 5: the oracle's literal restatement on one thread
        if args.cpu_max and n <= args.cpu_max:
            from oracle import binding as ob
            t0 = time.perf_counter()
            feats = ob.ppf_estimation(model, n_threads=1)
            t1 = time.perf_counter()
            hm = ob.HashMap(workloads.ANGLE_STEP, workloads.DIST_STEP).set_input_feature_cloud(feats, n_threads=1)
            t2 = time.perf_counter()
            assert hm.num_entries == best["entries"] or abs(hm.num_entries - best["entries"]) <= 64
            cpu_ref = (n * (n - 1), t1 - t0, t2 - t1)
            best["cpu"] = {"estimation_s": t1 - t0, "hashmap_s": t2 - t1, "threads": 1, "extrapolated": False,
                           "note": "oracle port; above 1024 model points its container mixes the key hash (PCL's own hash makes the "
                                   "build quadratic: 41 min for 2 009 points), i.e. this CPU side is faster than PCL as shipped"}
            del hm, feats
        elif cpu_ref is not None:
            k = n * (n - 1) / cpu_ref[0]
            best["cpu"] = {"estimation_s": cpu_ref[1] * k, "hashmap_s": cpu_ref[2] * k, "threads": 1, "extrapolated": True,
                           "note": "scaled with the pairs from the largest measured build (SURVEY.md §8d: the multimap of 50 000 points "
                                   "would need ~150 GB)"}
        if "cpu" in best:
            best["speedup_vs_cpu"] = (best["cpu"]["estimation_s"] + best["cpu"]["hashmap_s"]) / (best["device_ms"] * 1e-3)
        print(json.dumps(best), flush=True)
        dm.free()


if __name__ == "__main__":
    main()
