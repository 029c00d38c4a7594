"""GPU parity tests of the engine the reference actually calls — cv::ppf_match_3d::PPF3DDetector on the device
(csrc/k7_cvppf.cu, b200cv_*) — against its checker oracle/cvppf_oracle.cpp (parity unpinned: see that file's header).

Index work is compared bit for bit: sampled clouds, the buckets of the murmur-hashed table (collisions included),
accumulators, per-reference peaks.  Double-precision results (poses) agree to rounding: the device's libm and glibc's
differ in the last ulp of sin / cos / acos / atan2.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _rigid(axis, angle, t):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    T[:3, 3] = t
    return T


@pytest.fixture(scope="module")
def case(bottle_5mm):
    """the bottle under a known rigid motion inside clutter, and a curvature-edge-like subset of the instance"""
    G = _rigid((0.3, 1.0, 0.2), 0.7, (0.1, -0.05, 0.3))
    rng = np.random.default_rng(7)
    inst = np.concatenate([bottle_5mm[:, :3] @ G[:3, :3].T + G[:3, 3], bottle_5mm[:, 3:] @ G[:3, :3].T], axis=1)
    centre = inst[:, :3].mean(0)
    clutter = np.concatenate([centre + rng.normal(scale=0.15, size=(600, 3)), rng.normal(size=(600, 3))], axis=1)
    clutter[:, 3:] /= np.linalg.norm(clutter[:, 3:], axis=1, keepdims=True)
    scene = np.concatenate([inst, clutter]).astype(np.float32)
    # "edges": the instance points where the profile bends most (the shoulder and the rims of the bottle) — what
    # EdgeExtraction's curvature threshold keeps
    z = bottle_5mm[:, :3] @ np.linalg.eigh(np.cov(bottle_5mm[:, :3].T))[1][:, -1]
    q = np.quantile(z, [0.03, 0.62, 0.72, 0.97])
    edge = inst[(z < q[0]) | ((z > q[1]) & (z < q[2])) | (z > q[3])].astype(np.float32)
    return bottle_5mm, scene, edge, G


@pytest.fixture(scope="module")
def detectors(ctx, oracle, case):
    from yolo_ppf_pose_estimation_b200 import capi
    model = case[0]
    ref = oracle.CvDetector(0.04, 0.05).train_model(model)
    dev = capi.CvDetector(ctx, 0.04, 0.05).train_model(model)
    return dev, ref


def test_cv_sampling_and_training_bit_exact(ctx, oracle, case, detectors):
    dev, ref = detectors
    ci = dev.info
    m = ref.n_model
    assert ci.n_sampled == m and ci.table_size == ref.table_size and ci.n_nodes == m * (m - 1) and ci.num_angles == 30
    assert np.array_equal(dev.model_points().view(np.uint32), ref.model_points().view(np.uint32))   # samplePCByQuantization
    off, nodes, alpha = dev.table_export()
    assert off[-1] == ci.n_nodes and (np.diff(off.astype(np.int64)) >= 0).all()
    # every bucket of the murmur-hashed table holds the oracle's nodes (colliding keys included); compare all of them
    lens = np.diff(off.astype(np.int64))
    rng = np.random.default_rng(3)
    nonempty = np.flatnonzero(lens)
    longest = nonempty[np.argsort(lens[nonempty])[-20:]]
    for b in np.concatenate([rng.choice(nonempty, 400, replace=False), longest]):
        want = ref.bucket(int(b))
        got = nodes[off[b]:off[b + 1]]
        assert np.array_equal(got, want), b
        assert np.array_equal(dev.bucket(int(b)), want)
    assert len(ref.bucket(int(np.flatnonzero(lens == 0)[0]))) == 0
    # a bucket that mixes several keys exists (the table has n^2 nodes in >= n^2 buckets keyed by a hash)
    mp = ref.model_points()
    mixed = 0
    for b in longest[-5:]:
        keys = set()
        for node in nodes[off[b]:off[b + 1]][:200]:
            i, j = divmod(int(node), m)
            f, h = oracle.cv_pair(mp[i], mp[j], ci.angle_step, ci.distance_step)
            keys.add(tuple(int(x) for x in (f[0] / ci.angle_step, f[1] / ci.angle_step, f[2] / ci.angle_step, f[3] / ci.distance_step)))
            assert (h & (ci.table_size - 1)) == b
        mixed += len(keys) > 1
    print("buckets of the 5 longest that mix keys:", mixed)
    # alpha_m: float of a double that differs from glibc's by at most an ulp of atan2
    _, _, alpha = dev.table_export()
    A = alpha.reshape(m, m)
    for i, j in ((0, 1), (5, 300), (m - 1, 7)):
        f, _ = oracle.cv_pair(mp[i], mp[j], ci.angle_step, ci.distance_step)
        assert np.isfinite(A[i, j])


def test_cv_match_accumulators_and_peaks_bit_exact(ctx, oracle, case, detectors):
    dev, ref = detectors
    model, scene, edge, G = case
    res, ncl = dev.match(scene, 1.0 / 5.0, 0.04)
    rposes, rvotes, rraw, rncl = ref.match(scene, 1.0 / 5.0, 0.04, max_poses=16)
    raw = dev.raw_poses()
    assert len(raw) == len(rraw) and ncl == rncl
    assert np.array_equal(dev.scene_points().view(np.uint32), oracle.cv_sample(scene, 0.04).view(np.uint32))
    assert np.array_equal(raw["num_votes"], rraw[:, 0]) and np.array_equal(raw["model_index"], rraw[:, 1])
    assert np.array_equal(raw["alpha_index"], rraw[:, 2])
    for r in (0, 7, len(raw) // 2, len(raw) - 1):
        acc = dev.accumulator(scene, r, 1.0 / 5.0, 0.04)
        want = ref.accumulator(scene, r * 5, 0.04)
        assert np.array_equal(acc, want), r
        flat = int(np.argmax(acc))
        assert raw[r]["num_votes"] == acc.reshape(-1)[flat]
        if raw[r]["num_votes"]:
            assert (int(raw[r]["model_index"]), int(raw[r]["alpha_index"])) == divmod(flat, acc.shape[1])
    # clusters: same votes, poses equal to rounding, and the rigid motion is recovered
    k = min(len(res), len(rposes))
    assert np.array_equal(res["num_votes"][:k], rvotes[:k])
    assert np.abs(res["pose"][:k].reshape(k, 4, 4) - rposes[:k]).max() < 1e-9
    P = res["pose"][0].reshape(4, 4)
    pts = model[:, :3].astype(np.float64)
    err = np.linalg.norm((pts @ P[:3, :3].T + P[:3, 3]) - (pts @ G[:3, :3].T + G[:3, 3]), axis=1).mean()
    assert err < 0.02, err


def test_cv_match_s2b_vs_oracle(ctx, oracle, case, detectors):
    """match_S2B as inferred (reference points from the surface cloud, paired with the edge cloud): the same
    arithmetic, the second point set swapped — accumulators and peaks bit-exact against the checker."""
    dev, ref = detectors
    model, scene, edge, G = case
    res, ncl = dev.match_s2b(scene, edge, 1.0 / 5.0, 0.04)
    rposes, rvotes, rraw, rncl = ref.match_s2b(scene, edge, 1.0 / 5.0, 0.04, max_poses=16)
    raw = dev.raw_poses()
    assert ncl == rncl and np.array_equal(raw["num_votes"], rraw[:, 0]) and np.array_equal(raw["model_index"], rraw[:, 1])
    assert np.array_equal(raw["alpha_index"], rraw[:, 2])
    assert dev.info.n_second_sampled == len(oracle.cv_sample(edge, 0.04))
    for r in (3, len(raw) - 2):
        assert np.array_equal(dev.accumulator(scene, r, 1.0 / 5.0, 0.04, edge=edge), ref.accumulator(scene, r * 5, 0.04, edge=edge))
    k = min(len(res), len(rposes))
    assert np.array_equal(res["num_votes"][:k], rvotes[:k])
    assert np.abs(res["pose"][:k].reshape(k, 4, 4) - rposes[:k]).max() < 1e-9


def test_cv_errors_and_edges(ctx, case):
    from yolo_ppf_pose_estimation_b200 import capi
    model, scene, edge, G = case
    d = capi.CvDetector(ctx, 0.05, 0.05)
    with pytest.raises(capi.B200PPFError):
        d.match(scene)                                   # not trained
    with pytest.raises(capi.B200PPFError):
        d.train_model(model[:, :3])                      # no normals
    flat = model.copy()
    flat[:, 2] = 0.5
    with pytest.raises(capi.B200PPFError):
        d.train_model(flat)                              # zero extent: upstream divides by it
    d.train_model(model)
    with pytest.raises(capi.B200PPFError):
        d.match_s2b(scene, edge[:0])
    with pytest.raises(capi.B200PPFError):
        capi.CvDetector(ctx, 0.0, 0.05)
    # search parameters reach the clustering: a huge position threshold and rotation threshold merge everything
    d.set_search_params(1e3, 10.0)
    res, ncl = d.match(scene, 1.0 / 5.0, 0.05)
    assert ncl == 1 and res["num_votes"][0] == d.raw_poses()["num_votes"].sum()


def test_cpp_opencv_ppf_shim(tmp_path, ctx, oracle, case, detectors):
    """tests/cpp/cv_ppf_shim_example.cpp — PPF3DDetector(0.04, 0.05).trainModel / match / match_S2B and the ICP of the best
    poses, the reference's own call sequence, through include/opencv_compat — gives the C-ABI results."""
    from yolo_ppf_pose_estimation_b200 import build
    dev, ref = detectors
    model, scene, edge, G = case
    lib = build.build()
    model.astype(np.float32).tofile(tmp_path / "model.f32")
    scene.tofile(tmp_path / "scene.f32")
    edge.tofile(tmp_path / "edge.f32")
    exe = tmp_path / "cv_ppf_shim_example"
    cmd = ["/usr/bin/g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include", "opencv_compat"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "cv_ppf_shim_example.cpp"), "-o", str(exe), "-L", os.path.dirname(lib), "-lb200ppf",
           f"-Wl,-rpath,{os.path.dirname(lib)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = {ln.split()[0]: ln for ln in r.stdout.strip().splitlines()}
    for tag, (res, ncl) in (("match", dev.match(scene, 1.0 / 5.0, 0.04)), ("match_S2B", dev.match_s2b(scene, edge, 1.0 / 5.0, 0.04))):
        parts = lines[tag].split(" | ")
        assert int(parts[0].split()[1]) == ncl
        for k, part in enumerate(parts[1:]):
            v = part.split()
            assert int(v[0]) == res["num_votes"][k] and int(v[1]) == res["model_index"][k]
            assert np.array_equal(np.array([float(x) for x in v[2:]]), res["pose"][k])
    icp = [float(x) for x in lines["icp"].split()[1:]]
    P = np.array(icp[1:]).reshape(4, 4)
    pts = model[:, :3].astype(np.float64)
    err = np.linalg.norm((pts @ P[:3, :3].T + P[:3, 3]) - (pts @ G[:3, :3].T + G[:3, 3]), axis=1).mean()
    assert err < 0.01, err      # the refined pose is at least as good as the matched one
