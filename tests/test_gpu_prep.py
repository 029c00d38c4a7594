"""GPU parity tests of the scene pre-processing stages (prep.cu) through the C ABI, against the CPU oracle
(oracle/prep_oracle.cpp) on the reference's own crop (tests/golden/scene_crop_raw.npz).

Bars: everything that is index, byte or un-fused float +,-,*,/,sqrt work is bit-exact (voxel centroids, neighbour
lists and squared distances, mean neighbour distances, kept indices, re-normalised normals); the outlier
threshold is a double-precision sum over the cloud formed in a different (fixed) order: 1e-12 relative, and a
point may flip only inside that band; normals and curvature go through atan2f / cosf / sinf, whose device and
glibc versions differ by ulps: 2e-4 per component / 2e-5 absolute.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ANGLE_STEP, DIST_STEP, ROOT, load_cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def crop_raw():
    return load_cloud("scene_crop_raw")


@pytest.fixture(scope="module")
def dev_raw(ctx, crop_raw):
    return ctx.upload_xyz(crop_raw)


@pytest.fixture(scope="module")
def crop_5mm(oracle, crop_raw):
    return oracle.voxel_grid(crop_raw, 0.005)[0]


def test_upload_xyz_download_round_trip(ctx, crop_raw, dev_raw):
    assert dev_raw.size == crop_raw.shape[0]
    out = dev_raw.download(curvature=True)
    assert np.array_equal(out[:, :3], crop_raw) and not out[:, 3:].any()
    # non-finite points are dropped, the rest keep their order
    bad = crop_raw[:100].copy()
    bad[7, 1] = np.nan
    bad[50, 2] = np.inf
    c = ctx.upload_xyz(bad)
    assert c.size == 98 and np.array_equal(c.download()[:, :3], np.delete(bad, [7, 50], axis=0))
    assert ctx.upload_xyz(np.zeros((0, 3), np.float32)).size == 0


def _crop_box():
    b = np.load(os.path.join(ROOT, "tests", "golden", "crop_box.npz"))
    depth = np.zeros(tuple(b["image"]), np.float32)
    for (r, c), d in zip(b["pixels"], b["depths"]):
        depth[r, c] = d
    return depth, b["box"], b["intrinsics"]


def test_p0_frustum_crop_vs_oracle(ctx, oracle, scene_full):
    """SceneCropping for the surrogate YOLO box: corners on the host, ConvexHull + CropHull as five half-spaces"""
    from yolo_ppf_pose_estimation_b200 import capi
    depth, box, K = _crop_box()
    cor = capi.frustum_corners(depth, box, K)
    assert np.array_equal(cor, oracle.frustum_corners(depth, box, K))
    rng = np.random.default_rng(5)
    cloud = np.concatenate([scene_full, (rng.random((50000, 6)) * [1.0, 1.2, 2.0, 1, 1, 1] - [0.6, 0.7, 0.2, 0, 0, 0]).astype(np.float32)])
    keep = oracle.crop_pyramid(cloud[:, :3], cor)
    out, kept = ctx.crop_pyramid(ctx.upload_cloud(cloud), cor)
    assert np.array_equal(kept, np.flatnonzero(keep))  # no point of these sits within rounding of a face
    assert np.array_equal(out.download(), cloud[keep])  # order and normals kept
    # nothing inside / everything inside / empty input
    assert ctx.crop_pyramid(ctx.upload_xyz(cloud[:, :3] + np.float32(10.0)), cor)[0].size == 0
    wide = cor * np.array([50, 50, 10], np.float32)
    assert ctx.crop_pyramid(ctx.upload_xyz(cloud[keep][:, :3]), wide)[0].size == int(keep.sum())
    assert ctx.crop_pyramid(ctx.upload_xyz(np.zeros((0, 3), np.float32)), cor)[0].size == 0
    bent = cor.copy()
    bent[3, 2] += 0.1  # the hull of a non-planar base is another solid
    with pytest.raises(capi.B200PPFError):
        ctx.crop_pyramid(ctx.upload_xyz(cloud[:, :3]), bent)


@pytest.mark.parametrize("leaf", [0.01, 0.005, (0.02, 0.01, 0.005)])
def test_p1_voxel_grid_vs_oracle(ctx, oracle, crop_raw, dev_raw, leaf):
    ref, overflow = oracle.voxel_grid(crop_raw, leaf)
    out = ctx.voxel_grid(dev_raw, leaf)
    assert not overflow and out.size == ref.shape[0]
    got = out.download()
    assert np.array_equal(got[:, :3], ref)  # same voxels, same order, same float sums
    assert not got[:, 3:].any()
    assert ctx.timings()["prep_ms"] > 0


def test_p1_voxel_grid_edge_cases(ctx, oracle):
    assert ctx.voxel_grid(ctx.upload_xyz(np.zeros((0, 3), np.float32)), 0.01).size == 0
    one = np.array([[0.1, -0.2, 0.7]], np.float32)
    assert np.array_equal(ctx.voxel_grid(ctx.upload_xyz(one), 0.01).download()[:, :3], one)
    # every point in one leaf
    rng = np.random.default_rng(1)
    blob = (rng.random((5000, 3)) * 0.009 + 0.0005).astype(np.float32)
    got = ctx.voxel_grid(ctx.upload_xyz(blob), 0.01).download()[:, :3]
    assert np.array_equal(got, oracle.voxel_grid(blob, 0.01)[0]) and got.shape[0] == 1
    # negative coordinates, many leaves
    cloud = ((rng.random((20000, 3)) - 0.5) * np.array([2.0, 1.0, 0.5])).astype(np.float32)
    assert np.array_equal(ctx.voxel_grid(ctx.upload_xyz(cloud), 0.03).download()[:, :3], oracle.voxel_grid(cloud, 0.03)[0])
    # PCL: leaf too small for the extent -> warning, input returned unchanged
    far = np.array([[0, 0, 0], [100, 100, 100]], np.float32)
    assert np.array_equal(ctx.voxel_grid(ctx.upload_xyz(far), 0.001).download()[:, :3], far)
    from yolo_ppf_pose_estimation_b200 import capi
    with pytest.raises(capi.B200PPFError):
        ctx.voxel_grid(ctx.upload_xyz(far), 0.0)


@pytest.mark.parametrize("k", [8, 31, 51, 100])
def test_p2_knn_vs_oracle(ctx, oracle, crop_5mm, k):
    idx, d2 = oracle.knn(crop_5mm, k)
    gi, gd = ctx.knn(ctx.upload_xyz(crop_5mm), k)
    assert np.array_equal(gd, d2)
    assert np.array_equal(gi, idx)


def test_p2_knn_ties_and_the_full_scene(ctx, oracle, scene_full):
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3)
    g = np.concatenate([g, g[:100]]).astype(np.float32) * np.float32(0.01)
    for k in (8, 64):
        idx, d2 = oracle.knn(g, k)
        gi, gd = ctx.knn(ctx.upload_xyz(g), k)
        assert np.array_equal(gi, idx) and np.array_equal(gd, d2)
    xyz = np.ascontiguousarray(scene_full[:, :3])  # 44 893 points, several surfaces and depths
    idx, d2 = oracle.knn(xyz, 31)
    gi, gd = ctx.knn(ctx.upload_xyz(xyz), 31)
    assert np.array_equal(gd, d2) and np.array_equal(gi, idx)
    print("full-scene knn: prep_ms", ctx.timings()["prep_ms"])


@pytest.mark.parametrize("mean_k,mul", [(50, 1.0), (50, 1.5), (20, 0.5)])
def test_p3_outlier_removal_vs_oracle(ctx, oracle, crop_5mm, mean_k, mul):
    keep, dist, thr = oracle.statistical_outlier_removal(crop_5mm, mean_k, mul)
    out, kept, gdist, gthr = ctx.statistical_outlier_removal(ctx.upload_xyz(crop_5mm), mean_k, mul)
    assert np.array_equal(gdist, dist)  # float sqrt + double sums in neighbour order: identical
    assert abs(gthr - thr) <= 1e-12 * abs(thr)
    gkeep = np.zeros(crop_5mm.shape[0], bool)
    gkeep[kept] = True
    undecided = np.abs(dist.astype(np.float64) - thr) <= 1e-12 * abs(thr)
    assert np.array_equal(gkeep[~undecided], keep[~undecided])
    assert np.all(np.diff(kept.astype(np.int64)) > 0)
    assert np.array_equal(out.download()[:, :3], crop_5mm[kept])
    from yolo_ppf_pose_estimation_b200 import capi
    with pytest.raises(capi.B200PPFError):  # PCL would read past its neighbour list
        ctx.statistical_outlier_removal(ctx.upload_xyz(crop_5mm[:mean_k]), mean_k, mul)


@pytest.mark.parametrize("cov_mode", [0, 1])
def test_p4_normals_vs_oracle(ctx, oracle, crop_5mm, cov_mode):
    ref = oracle.normals(crop_5mm, 30, cov_mode=cov_mode)
    cloud = ctx.upload_xyz(crop_5mm)
    ctx.normal_estimation(cloud, 30, covariance_mode=cov_mode)
    got = cloud.download(curvature=True)
    assert np.array_equal(got[:, :3], crop_5mm)
    dn = np.abs(got[:, 3:6] - ref[:, :3]).max()
    dc = np.abs(got[:, 6] - ref[:, 3]).max()
    print(f"normals cov_mode {cov_mode}: max component diff {dn:.3g}, max curvature diff {dc:.3g}")
    assert dn < 2e-4 and dc < 2e-5
    assert np.all(np.einsum("ni,ni->n", got[:, 3:6], -crop_5mm) >= 0)  # flipped towards the camera at the origin
    # another viewpoint, another k; fewer than three points -> NaN as PCL writes them
    vp = (0.3, -0.1, 2.0)
    ctx.normal_estimation(cloud, 12, viewpoint=vp, covariance_mode=cov_mode)
    got = cloud.download(curvature=True)
    ref = oracle.normals(crop_5mm, 12, viewpoint=vp, cov_mode=cov_mode)
    assert np.abs(got[:, 3:6] - ref[:, :3]).max() < 2e-4 and np.abs(got[:, 6] - ref[:, 3]).max() < 2e-5
    two = ctx.upload_xyz(crop_5mm[:2])
    ctx.normal_estimation(two, 30)
    assert np.isnan(two.download()[:, 3:]).all()


def test_p5_edges_and_p6_renormalise(ctx, oracle, crop_5mm):
    cloud = ctx.upload_xyz(crop_5mm)
    ctx.normal_estimation(cloud, 30)
    full = cloud.download(curvature=True)
    thr = np.float32(0.03)  # EdgeExtraction(0.03), src/YOLO_cropping_ppf_test.cpp:103
    edges = ctx.curvature_edges(cloud, thr).download(curvature=True)
    sel = full[:, 6] > thr
    assert 0 < sel.sum() < full.shape[0]
    assert np.array_equal(edges, full[sel])
    assert ctx.curvature_edges(cloud, 10.0).size == 0
    ctx.normalize_normals(cloud)
    got = cloud.download()
    assert np.array_equal(got[:, 3:], oracle.renormalize_normals(full[:, 3:6]))


def test_prep_chain_feeds_the_ppf_engine(ctx, oracle, crop_raw, bottle):
    """raw crop -> Subsampling(1 cm) -> OutlierProcessing(50, 1.0) -> NormalEstimation(30) -> N x 6 -> align, all on the
    device, against the same chain on the CPU (src/YOLO_cropping_ppf_test.cpp:96-122)"""
    d = ctx.voxel_grid(ctx.upload_xyz(crop_raw), 0.01)
    d, kept, _, _ = ctx.statistical_outlier_removal(d, 50, 1.0)
    ctx.normal_estimation(d, 30)
    ctx.normalize_normals(d)
    scene_gpu = d.download()

    v = oracle.voxel_grid(crop_raw, 0.01)[0]
    keep, _, _ = oracle.statistical_outlier_removal(v, 50, 1.0)
    v = v[keep]
    n = oracle.normals(v, 30)
    scene_cpu = np.concatenate([v, oracle.renormalize_normals(n[:, :3])], axis=1)
    assert np.array_equal(scene_gpu[:, :3], scene_cpu[:, :3])
    assert np.abs(scene_gpu[:, 3:] - scene_cpu[:, 3:]).max() < 2e-4

    # the device cloud goes straight into the voting kernels (its nrm.w holds the curvature: never read there)
    model = ctx.upload_cloud(bottle)
    table = ctx.table_build_from_cloud(model, ANGLE_STEP, DIST_STEP)
    final, poses, votes = ctx.register(model, table, d, ref_rate=5)
    final_up, _, votes_up = ctx.register(model, table, ctx.upload_cloud(scene_gpu), ref_rate=5)
    assert np.array_equal(final, final_up) and np.array_equal(votes, votes_up)
    hm = oracle.HashMap(ANGLE_STEP, DIST_STEP).set_input_feature_cloud(oracle.ppf_estimation(bottle))
    rfinal, _, rvotes, _ = hm.register(bottle, scene_gpu, ref_rate=5, n_threads=oracle.max_threads())
    dt = float(np.linalg.norm(final[:3, 3] - rfinal[:3, 3]))
    R = final[:3, :3].astype(np.float64).T @ rfinal[:3, :3].astype(np.float64)
    dr = float(np.degrees(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))))
    print(f"chain: {scene_gpu.shape[0]} scene points, votes {votes.tolist()} vs {rvotes.tolist()}, |dt| {dt:.2e} m, dR {dr:.3f} deg")
    assert dt < 1e-3 and dr < 0.5


def test_cpp_pcl_prep_shim_end_to_end(tmp_path, ctx, crop_raw):
    """tests/cpp/pcl_prep_example.cpp — VoxelGrid -> StatisticalOutlierRemoval -> NormalEstimationOMP ->
    concatenateFields -> edges -> N x 6 through include/pcl_compat — gives the C-ABI chain's result."""
    from yolo_ppf_pose_estimation_b200 import build
    lib = build.build()
    crop_raw.astype(np.float32).tofile(tmp_path / "scene_crop_raw.f32")
    exe = tmp_path / "pcl_prep_example"
    cmd = ["/usr/bin/g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include", "pcl_compat"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "pcl_prep_example.cpp"), "-o", str(exe),
           "-L", os.path.dirname(lib), "-lb200ppf", f"-Wl,-rpath,{os.path.dirname(lib)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path / "scene_crop_raw.f32"), "0.005", "1.0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    d = ctx.voxel_grid(ctx.upload_xyz(crop_raw), 0.005)
    n_voxel = d.size
    d, _, _, _ = ctx.statistical_outlier_removal(d, 50, 1.0)
    ctx.normal_estimation(d, 30)
    n_edges = ctx.curvature_edges(d, 0.03).size
    ctx.normalize_normals(d)
    mat = d.download()
    assert [int(x) for x in lines[0].split()[1:]] == [crop_raw.shape[0], n_voxel, mat.shape[0], mat.shape[0], n_edges]
    for i in range(3):
        assert np.array_equal(np.array([float(x) for x in lines[1 + i].split()[1:]], np.float32), mat[i])
    assert float(lines[4].split()[1]) == float(np.sum(mat.astype(np.float64).reshape(-1).cumsum()[-1:]))


def test_match_object_equals_the_stepwise_chain(ctx, oracle, scene_full, bottle):
    """b200ppf_match_object = SceneCropping -> Subsampling -> OutlierProcessing -> NormalEstimation -> EdgeExtraction ->
    re-normalise -> align -> ICP in one call (src/YOLO_cropping_ppf_test.cpp:91-122): the same result as the stages
    called one by one, and the pose the CPU chain finds on the device's own object cloud"""
    from yolo_ppf_pose_estimation_b200 import capi
    depth, box, K = _crop_box()
    cor = capi.frustum_corners(depth, box, K)
    scene = ctx.upload_xyz(scene_full[:, :3])
    model = ctx.upload_cloud(bottle)
    table = ctx.table_build_from_cloud(model, ANGLE_STEP, DIST_STEP)
    res, obj, edges = ctx.match_object(scene, cor, model, table, leaf=0.01, ref_rate=5, icp_max_iterations=30, icp_num_levels=4)
    # stage by stage
    d, _ = ctx.crop_pyramid(scene, cor)
    n_crop = d.size
    d = ctx.voxel_grid(d, 0.01)
    n_vox = d.size
    d, _, _, _ = ctx.statistical_outlier_removal(d, 50, 1.0)
    ctx.normal_estimation(d, 30)
    e = ctx.curvature_edges(d, 0.03)
    ctx.normalize_normals(d)
    ctx.normalize_normals(e)
    assert (res["n_cropped"], res["n_sampled"], res["n_filtered"], res["n_edges"]) == (n_crop, n_vox, d.size, e.size)
    assert np.array_equal(obj.download(curvature=True), d.download(curvature=True))
    assert np.array_equal(edges.download(curvature=True), e.download(curvature=True))
    final, poses, votes = ctx.register(model, table, d, ref_rate=5)
    P, resid, _ = ctx.icp_refine(model, d, poses.astype(np.float64), max_iterations=30, num_levels=4)
    assert res["n_poses"] == len(poses) and res["votes"] == votes[0]
    assert np.array_equal(res["pose"], P[0]) and res["residual"] == resid[0]
    assert res["total_wall_ms"] > 0 and res["match_ms"] > 0 and res["icp_ms"] > 0
    # the CPU chain on the same object cloud
    obj_host = d.download()
    hm = oracle.HashMap(ANGLE_STEP, DIST_STEP).set_input_feature_cloud(oracle.ppf_estimation(bottle))
    _, rposes, rvotes, _ = hm.register(bottle, obj_host, ref_rate=5, n_threads=oracle.max_threads())
    R, _, _ = oracle.icp_refine(bottle, obj_host, rposes[:1].astype(np.float64), max_iterations=30, num_levels=4)
    dt = float(np.linalg.norm(res["pose"][:3, 3] - R[0][:3, 3]))
    print(f"object: {n_crop} cropped -> {n_vox} -> {d.size} points, votes {res['votes']} vs {rvotes[0]}, |dt| vs CPU chain {dt:.2e} m, "
          f"wall {res['total_wall_ms']:.2f} ms")
    assert dt < 1e-3
    with pytest.raises(capi.B200PPFError):  # nothing inside the frustum
        ctx.match_object(ctx.upload_xyz(scene_full[:, :3] + np.float32(10)), cor, model, table)


def test_full_size_1Mi_scene_equals_the_host_build_of_the_kernels(ctx, oracle):
    """BASELINE config 4's frame (1 Mi points -> 684 438 after the 5 mm voxel grid): every stage against the checker
    (voxel grid) or against the kernels' own query code run on the CPU (b200ppf_debug_knn_host) — which tests/test_prep.py
    ties to the checker at small sizes and to a kd-tree at this size.  Index and float work bit-exact, normals within the
    libm band."""
    from yolo_ppf_pose_estimation_b200 import capi, synth
    scene = np.ascontiguousarray(synth.synth_library_scene(1 << 20)[:, :3], np.float32)
    v = ctx.voxel_grid(ctx.upload_xyz(scene), 0.005)
    vh = v.download()[:, :3]
    assert np.array_equal(vh, oracle.voxel_grid(scene, 0.005)[0]) and v.size == 684438
    gi, gd = ctx.knn(v, 51)
    hi, hd = capi.debug_knn_host(vh, 51, 0)
    assert np.array_equal(gd, hd) and np.array_equal(gi, hi)
    out, kept, gdist, thr = ctx.statistical_outlier_removal(v, 50, 1.0)
    # the epilogue of the same search: sqrtf of the 50 other neighbours summed in double, in order
    mean = (np.sqrt(hd[:, 1:]).astype(np.float64).cumsum(axis=1)[:, -1] / 50).astype(np.float32)
    assert np.array_equal(gdist, mean)
    m64 = mean.astype(np.float64)
    assert abs(thr - (m64.mean() + m64.std(ddof=1))) <= 1e-9 * thr
    assert np.array_equal(kept, np.flatnonzero(~(m64 > thr))) and out.size == 540433
    ctx.normal_estimation(out, 30)
    gn = out.download(curvature=True)
    hn = capi.debug_knn_host(gn[:, :3], 30, 2)
    dn, dc = np.abs(gn[:, 3:6] - hn[:, :3]).max(1), np.abs(gn[:, 6] - hn[:, 3])
    print(f"1 Mi scene normals: max curvature diff {dc.max():.3g}, normal component diff p99.9 {np.quantile(dn, 0.999):.3g}, max {dn.max():.3g}")
    assert dc.max() < 2e-5 and (dn > 2e-4).mean() < 1e-3  # nearly isotropic neighbourhoods have no stable normal
    assert ctx.curvature_edges(out, 0.03).size in range(126399 - 20, 126399 + 21)
