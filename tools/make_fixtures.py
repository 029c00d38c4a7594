#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the reference's data files (run in the build container).

/root/reference does not exist on the GPU box, so everything the tests and bench.py need from
it is frozen here as small float32 fixtures:

  bottle_1cm.npz / bottle_5mm.npz   data/bottle_remesh_meter_normalized.ply voxel-averaged at
                                    1 cm / 5 mm (normals averaged + renormalised)          (C1/C2 model)
  scene_full_1cm.npz                data/1_depth.exr back-projected with the intrinsics of
                                    src/YOLO_cropping_ppf_test.cpp:35-37 (formula include/Camera.h:56-58;
                                    data/1_cloud.ply itself is missing, .MISSING_LARGE_BLOBS:1),
                                    voxel 1 cm, PCA normals k=30 flipped to the camera       (C2 scene)
  scene_crop_1cm.npz                the same cloud cropped with the frustum rule of
                                    include/CloudProcessing.h:279-332 around the surrogate YOLO box
                                    u in [536,631], v in [211,402]    (C1 scene)

  scene_crop_raw.npz                the crop before any filtering: raw back-projected points inside the
                                    frustum (xyz only) — the input of the pre-processing stages
                                    Subsampling / OutlierProcessing / NormalEstimation
                                    (include/CloudProcessing.h:340-401)
  crop_box.npz                      the surrogate box, the four depth pixels SceneCropping reads at its expanded
                                    corners and the intrinsics                 (`--raw-only` writes just these two)

Usage: OPENCV_IO_ENABLE_OPENEXR=1 python tools/make_fixtures.py [--raw-only]
"""
import os
import sys

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

import numpy as np
from scipy.spatial import cKDTree

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

FX, FY, CX, CY = 614.384, 614.365, 638.121, 364.01  # src/YOLO_cropping_ppf_test.cpp:35-37


def read_ply_ascii(path):
    with open(path) as f:
        n = 0
        while True:
            line = f.readline().strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        data = np.loadtxt(f, dtype=np.float64, max_rows=n)
    return data[:, :6].astype(np.float32)


def voxel_average(cloud, leaf):
    """PCL VoxelGrid semantics: centroid of every occupied leaf; normals averaged, renormalised."""
    xyz = cloud[:, :3].astype(np.float64)
    idx = np.floor(xyz / leaf).astype(np.int64)
    idx -= idx.min(axis=0)
    dims = idx.max(axis=0) + 1
    lin = (idx[:, 2] * dims[1] + idx[:, 1]) * dims[0] + idx[:, 0]
    order = np.argsort(lin, kind="stable")
    lin_s = lin[order]
    starts = np.flatnonzero(np.r_[True, lin_s[1:] != lin_s[:-1]])
    counts = np.diff(np.r_[starts, lin_s.size])
    sums = np.add.reduceat(cloud[order].astype(np.float64), starts, axis=0)
    out = sums / counts[:, None]
    if cloud.shape[1] >= 6:
        nrm = out[:, 3:6]
        ln = np.linalg.norm(nrm, axis=1, keepdims=True)
        out[:, 3:6] = nrm / np.where(ln > 1e-12, ln, 1.0)
    return out.astype(np.float32)


def pca_normals(xyz, k=30):
    """NormalEstimation(k) restated: smallest-eigenvector of the k-NN covariance, flipped to the origin."""
    tree = cKDTree(xyz)
    _, nn = tree.query(xyz, k=k)
    pts = xyz[nn].astype(np.float64)  # (n, k, 3)
    mu = pts.mean(axis=1, keepdims=True)
    d = pts - mu
    cov = np.einsum("nki,nkj->nij", d, d) / k
    w, v = np.linalg.eigh(cov)
    n = v[:, :, 0]
    flip = np.einsum("ni,ni->n", n, -xyz.astype(np.float64)) < 0
    n[flip] *= -1
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return n.astype(np.float32)


def backproject(depth):
    v, u = np.nonzero(depth > 0)
    z = depth[v, u].astype(np.float64)
    x = (u - CX) * z / FX
    y = (v - CY) * z / FY
    return np.stack([x, y, z], axis=1).astype(np.float32), u, v


def frustum_crop(xyz, depth, box, margin_px=30, margin_z=0.15):
    """include/CloudProcessing.h:279-332: four margin-expanded corner rays at the mean corner depth."""
    u0, v0, u1, v1 = box
    u0, v0, u1, v1 = u0 - margin_px, v0 - margin_px, u1 + margin_px, v1 + margin_px
    h, w = depth.shape
    corners = [(u0, v0), (u1, v0), (u1, v1), (u0, v1)]
    ds = [float(depth[int(np.clip(v, 0, h - 1)), int(np.clip(u, 0, w - 1))]) for (u, v) in corners]
    zavg = float(np.mean(ds))
    zmax = zavg + margin_z
    # inside test in pixel space (equivalent to the hull of {origin, 4 corner rays}) and z <= zmax
    z = xyz[:, 2]
    u = xyz[:, 0] * FX / np.maximum(z, 1e-9) + CX
    v = xyz[:, 1] * FY / np.maximum(z, 1e-9) + CY
    return (u >= u0) & (u <= u1) & (v >= v0) & (v <= v1) & (z <= zmax) & (z > 0)


def save_crop_box(depth, box=(536, 211, 95, 191)):
    """crop_box.npz: the surrogate YOLO box (x, y, width, height) and the four depth pixels SceneCropping reads at
    its 30-pixel-expanded corners (include/CloudProcessing.h:279-288) — all the crop needs from the depth image"""
    x, y, w, h = box
    rows, cols = depth.shape
    left, top = max(x - 30, 0), max(y - 30, 0)
    right = x + w + 30 if x + w + 30 < cols else cols - 1
    bottom = y + h + 30 if y + h + 30 < rows else rows - 1
    px = np.array([[top, left], [top, right], [bottom, left], [bottom, right]], np.int32)  # (row, col) of depth_1..4
    np.savez_compressed(os.path.join(OUT, "crop_box.npz"), box=np.array(box, np.int32), image=np.array([rows, cols], np.int32),
                        pixels=px, depths=np.array([depth[r, c] for r, c in px], np.float32),
                        intrinsics=np.array([FX, FY, CX, CY], np.float64))


def main():
    import cv2

    os.makedirs(OUT, exist_ok=True)
    if "--raw-only" in sys.argv:
        depth = cv2.imread(os.path.join(REF, "data", "1_depth.exr"), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
        if depth.ndim == 3:
            depth = depth[:, :, 0]
        xyz, _, _ = backproject(depth)
        crop_raw = xyz[frustum_crop(xyz, depth, (536, 211, 631, 402))]
        print("scene_crop_raw", crop_raw.shape)
        np.savez_compressed(os.path.join(OUT, "scene_crop_raw.npz"), cloud=crop_raw.astype(np.float32))
        save_crop_box(depth)
        return 0
    bottle = read_ply_ascii(os.path.join(REF, "data", "bottle_remesh_meter_normalized.ply"))
    print("bottle", bottle.shape)
    for leaf, name in ((0.01, "bottle_1cm"), (0.005, "bottle_5mm")):
        m = voxel_average(bottle, leaf)
        print(name, m.shape)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), cloud=m)

    depth = cv2.imread(os.path.join(REF, "data", "1_depth.exr"), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_ANYCOLOR)
    if depth.ndim == 3:
        depth = depth[:, :, 0]
    xyz, _, _ = backproject(depth)
    print("scene raw", xyz.shape)
    full = voxel_average(xyz, 0.01)
    nrm = pca_normals(full[:, :3], 30)
    scene = np.concatenate([full[:, :3], nrm], axis=1).astype(np.float32)
    print("scene_full_1cm", scene.shape)
    np.savez_compressed(os.path.join(OUT, "scene_full_1cm.npz"), cloud=scene)

    keep = frustum_crop(xyz, depth, (536, 211, 631, 402))
    crop_raw = xyz[keep]
    crop = voxel_average(crop_raw, 0.01)
    nrm = pca_normals(crop[:, :3], 30)
    crop = np.concatenate([crop[:, :3], nrm], axis=1).astype(np.float32)
    print("scene_crop_1cm", crop_raw.shape, "->", crop.shape)
    np.savez_compressed(os.path.join(OUT, "scene_crop_1cm.npz"), cloud=crop)
    np.savez_compressed(os.path.join(OUT, "scene_crop_raw.npz"), cloud=crop_raw.astype(np.float32))
    save_crop_box(depth)


if __name__ == "__main__":
    sys.exit(main())
