#!/usr/bin/env python
"""Scene pre-processing timing (prep.cu): Subsampling -> OutlierProcessing(50, 1.0) -> NormalEstimation(30) ->
EdgeExtraction(0.03) -> re-normalise, as src/YOLO_cropping_ppf_test.cpp:96-121 chains them, B200 vs the CPU
restatement of the PCL operators on all host threads.  Two inputs: the reference's raw crop (29 450 points,
tests/golden/scene_crop_raw.npz) and a synthetic 1 Mi-point scene (Azure Kinect WFOV size, config C4).
Prints one JSON line per input.

usage: python tools/prep_bench.py [--no-cpu] [--repeat 5] [--cpu-points 60000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_chain(c, xyz, leaf, repeat):
    stages = {"upload": [], "voxel_grid": [], "outlier_removal": [], "normals": [], "edges": [], "renormalise": [], "total_wall": []}
    sizes = {}
    for _ in range(repeat + 1):
        t0 = time.perf_counter()
        d = c.upload_xyz(xyz)
        t1 = time.perf_counter()
        v = c.voxel_grid(d, leaf)
        stages["voxel_grid"].append(c.timings()["prep_ms"])
        f, kept, _, _ = c.statistical_outlier_removal(v, 50, 1.0)
        stages["outlier_removal"].append(c.timings()["prep_ms"])
        c.normal_estimation(f, 30)
        stages["normals"].append(c.timings()["prep_ms"])
        e = c.curvature_edges(f, 0.03)
        stages["edges"].append(c.timings()["prep_ms"])
        t2 = time.perf_counter()
        c.normalize_normals(f)
        c.synchronize()
        t3 = time.perf_counter()
        stages["upload"].append(1e3 * (t1 - t0))
        stages["renormalise"].append(1e3 * (t3 - t2))
        stages["total_wall"].append(1e3 * (t3 - t0))
        sizes = {"raw": int(xyz.shape[0]), "voxels": v.size, "kept": f.size, "edges": e.size}
    return {k: float(np.median(v[1:])) for k, v in stages.items()}, sizes


def cpu_chain(xyz, leaf):
    from oracle import binding as ob
    t = {}
    t0 = time.perf_counter()
    v = ob.voxel_grid(xyz, leaf)[0]
    t["voxel_grid"] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    keep, _, _ = ob.statistical_outlier_removal(v, 50, 1.0)
    t["outlier_removal"] = 1e3 * (time.perf_counter() - t0)
    v = v[keep]
    t0 = time.perf_counter()
    n = ob.normals(v, 30)
    t["normals"] = 1e3 * (time.perf_counter() - t0)
    t["threads"] = ob.max_threads()
    t["points_after_voxel"] = int(keep.shape[0])
    return t


def cpu_kdtree(c, xyz, leaf):
    """the neighbour searches of the two stages with a kd-tree on all host threads (scipy cKDTree standing in for
    the FLANN tree PCL builds), on the full-size voxelised cloud: the fair CPU figure for the search itself"""
    from scipy.spatial import cKDTree
    v = c.voxel_grid(c.upload_xyz(xyz), leaf).download()[:, :3].astype(np.float64)
    t0 = time.perf_counter()
    tree = cKDTree(v)
    t1 = time.perf_counter()
    tree.query(v, k=51, workers=-1)
    t2 = time.perf_counter()
    tree.query(v, k=30, workers=-1)
    t3 = time.perf_counter()
    return {"points": int(v.shape[0]), "build": 1e3 * (t1 - t0), "query_k51": 1e3 * (t2 - t1), "query_k30": 1e3 * (t3 - t2),
            "threads": os.cpu_count()}


def frame_case(name):
    """(scene xyz, corners, model N x 6, match_object overrides) for one detected object"""
    from yolo_ppf_pose_estimation_b200 import capi, synth
    gold = os.path.join(ROOT, "tests", "golden")
    if name == "reference frame":  # the reference's own data: regenerated scene (1 cm), surrogate YOLO box, bottle model
        b = np.load(os.path.join(gold, "crop_box.npz"))
        depth = np.zeros(tuple(b["image"]), np.float32)
        for (r, c), d in zip(b["pixels"], b["depths"]):
            depth[r, c] = d
        cor = capi.frustum_corners(depth, b["box"], b["intrinsics"])
        scene = np.load(os.path.join(gold, "scene_full_1cm.npz"))["cloud"][:, :3]
        model = np.load(os.path.join(gold, "bottle_1cm.npz"))["cloud"]
        return np.ascontiguousarray(scene, np.float32), cor, model.astype(np.float32), dict(leaf=0.01, ref_rate=5)
    # synthetic 1 Mi-point frame (config C4), library model 0: the box is the projection of its instance
    k = 0
    scene = np.ascontiguousarray(synth.synth_library_scene(1 << 20)[:, :3], np.float32)
    model = synth.synth_model(2000, 10 + k, k).astype(np.float32)
    T = synth.library_pose(k)
    p = model[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    zf = np.float32(p[:, 2].max() + 0.15)
    tx, ty = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
    xl, xr, yt, yb = (np.float32(v * zf) for v in (tx.min() - 0.03, tx.max() + 0.03, ty.min() - 0.03, ty.max() + 0.03))
    cor = np.array([[xl, yt, zf], [xl, yb, zf], [xr, yt, zf], [xr, yb, zf]], np.float32)
    return scene, cor, model, dict(leaf=0.005, ref_rate=20)  # 20 = the reference's 1 / 0.05


def frame_bench(c, name, repeat, with_cpu):
    """b200ppf_match_object (crop -> voxel -> outlier removal -> normals -> edges -> align -> ICP) per detected
    object, scene resident in HBM as it would be for every box of a frame, vs the same chain on the CPU"""
    from yolo_ppf_pose_estimation_b200 import capi
    scene, cor, model, over = frame_case(name)
    a = np.float32(12.0) / np.float32(180.0) * np.float32(np.pi)
    dm = c.upload_cloud(model)
    table = c.table_build_from_cloud(dm, a, np.float32(0.01))
    t0 = time.perf_counter()
    ds = c.upload_xyz(scene)
    upload_ms = 1e3 * (time.perf_counter() - t0)
    runs, obj_cloud = [], None
    for _ in range(repeat + 1):
        r, obj_cloud, _ = c.match_object(ds, cor, dm, table, **over)
        runs.append(r)
    runs = runs[1:]
    _, gpu_ppf_poses, _ = c.register(dm, table, obj_cloud, ref_rate=over["ref_rate"])  # the pose before ICP
    med = {k: float(np.median([r[k] for r in runs])) for k in runs[0] if k.endswith("_ms")}
    r = runs[-1]
    rec = {"frame": name, "scene_points": int(scene.shape[0]), "model_points": int(model.shape[0]), "params": over,
           "sizes": {k: int(r[k]) for k in ("n_cropped", "n_sampled", "n_filtered", "n_edges", "n_poses")}, "votes": int(r["votes"]),
           "scene_upload_ms": upload_ms, "b200_ms": med}
    if with_cpu:
        from oracle import binding as ob
        nt = ob.max_threads()
        hm = ob.HashMap(a, np.float32(0.01)).set_input_feature_cloud(ob.ppf_estimation(model))  # training is offline
        t = {}
        t0 = time.perf_counter()
        v = scene[ob.crop_pyramid(scene, cor)]
        t["crop"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        v = ob.voxel_grid(v, over["leaf"])[0]
        t["voxel"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        v = v[ob.statistical_outlier_removal(v, 50, 1.0)[0]]
        t["outlier"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        n = ob.normals(v, 30)
        obj = np.concatenate([v, ob.renormalize_normals(n[:, :3])], axis=1)
        t["normals"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        _, poses, votes, _ = hm.register(model, obj, ref_rate=over["ref_rate"], n_threads=nt)
        t["match"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        R, cres, _ = ob.icp_refine(model, obj, poses.astype(np.float64))
        t["icp"] = 1e3 * (time.perf_counter() - t0)
        t["total"] = float(sum(t.values()))
        rec["cpu_ms"] = t
        rec["cpu_threads"] = nt
        rec["cpu_note"] = "CPU restatement of the PCL / OpenCV operators, all host threads (ICP: one thread per pose, sequential); neighbour search by brute force"
        rec["ppf_translation_diff_vs_cpu_m"] = float(np.linalg.norm(gpu_ppf_poses[0][:3, 3] - poses[0][:3, 3]))
        # the reference's ICP reports 1e10 when it loses its correspondences (a wrong hypothesis): it is chaotic then,
        # and a comparison of the two diverged poses says nothing
        rec["icp_residual"] = {"b200": float(r["residual"]), "cpu": float(cres[0])}
        if r["residual"] < 1e9 and cres[0] < 1e9:
            rec["translation_diff_vs_cpu_m"] = float(np.linalg.norm(r["pose"][:3, 3] - R[0][:3, 3]))
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--only", choices=["crop", "big", "frames"], default=None, help="run one of the inputs / only the per-object frames")
    ap.add_argument("--no-frames", action="store_true", help="skip the per-object (b200ppf_match_object) lines")
    ap.add_argument("--cpu-points", type=int, default=60000, help="the CPU leg (brute-force neighbours) takes a raw subset that voxelises to about this many points")
    args = ap.parse_args()
    from yolo_ppf_pose_estimation_b200 import capi, synth
    c = capi.Context(0)
    crop = np.load(os.path.join(ROOT, "tests", "golden", "scene_crop_raw.npz"))["cloud"].astype(np.float32)
    big = np.ascontiguousarray(synth.synth_library_scene(1 << 20)[:, :3], np.float32)
    for name, xyz, leaf in (("reference crop", crop, 0.005), ("synthetic 1Mi scene", big, 0.005)):
        if args.only and (args.only == "frames" or (args.only == "crop") != (name == "reference crop")):
            continue
        ms, sizes = gpu_chain(c, xyz, leaf, args.repeat)
        rec = {"input": name, "leaf": leaf, "sizes": sizes, "b200_ms": ms,
               "b200_points_per_s": sizes["raw"] / (1e-3 * (ms["voxel_grid"] + ms["outlier_removal"] + ms["normals"] + ms["edges"]))}
        if not args.no_cpu:
            sub = xyz
            if sizes["voxels"] > args.cpu_points:  # the brute-force CPU neighbours are O(n^2): a spatial slab of the scene
                order = np.argsort(xyz[:, 0], kind="stable")
                sub = xyz[order[: int(xyz.shape[0] * args.cpu_points / sizes["voxels"])]]
            rec["cpu_ms"] = cpu_chain(sub, leaf)
            rec["cpu_sample_raw_points"] = int(sub.shape[0])
            rec["cpu_kdtree_ms"] = cpu_kdtree(c, xyz, leaf)
        print(json.dumps(rec), flush=True)
    if not args.no_frames and args.only in (None, "frames"):
        for name in ("reference frame", "synthetic 1Mi frame"):
            frame_bench(c, name, args.repeat, not args.no_cpu)


if __name__ == "__main__":
    main()
