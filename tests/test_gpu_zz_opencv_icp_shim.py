"""GPU test of the OpenCV-shaped ICP shim (include/opencv_compat): the reference's own refinement call,
ICP icp(100, 0.005f, 2.5f, 8); icp.registerModelToScene(model, scene, poses)  (include/CloudProcessing.h:518-523),
compiled against the shim and run on the B200, gives exactly what b200ppf_icp_refine gives through the C ABI
(which tests/test_gpu_parity.py::test_k6_icp_vs_oracle ties to the CPU restatement of opencv_contrib's icp.cpp)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_opencv_icp_shim_end_to_end(tmp_path, ctx):
    from yolo_ppf_pose_estimation_b200 import build
    from test_gpu_parity import _icp_case
    lib = build.build()
    model, scene, _, starts = _icp_case()
    model, scene = np.ascontiguousarray(model, np.float32), np.ascontiguousarray(scene, np.float32)
    starts = np.ascontiguousarray(starts, np.float64)
    model.tofile(tmp_path / "model.f32")
    scene.tofile(tmp_path / "scene.f32")
    starts.tofile(tmp_path / "poses.f64")
    exe = tmp_path / "cv_icp_shim_example"
    cmd = ["/usr/bin/g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include", "opencv_compat"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "cv_icp_shim_example.cpp"), "-o", str(exe),
           "-L", os.path.dirname(lib), "-lb200ppf", f"-Wl,-rpath,{os.path.dirname(lib)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(tmp_path / "refined.f64", np.float64).reshape(-1, 17)
    P, res, _ = ctx.icp_refine(ctx.upload_cloud(model), ctx.upload_cloud(scene), starts)
    assert out.shape[0] == len(starts) == 5
    assert np.array_equal(out[:, :16].reshape(-1, 4, 4), P) and np.array_equal(out[:, 16], res)
    assert "Pose to Model Index" in r.stdout
