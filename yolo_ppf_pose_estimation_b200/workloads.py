"""The workloads BASELINE.json names, as (model cloud, scene cloud, parameters).

  c1  bottle (1 cm voxel, 543 pts) vs the YOLO-cropped scene (934 pts), reference rate 5
  c2  the same bottle vs the full uncropped scene (44 893 pts), every scene point a reference   <- bench default
  c2_5mm  2 009-point bottle vs the same scene (needs two accumulator slices)
  c3  synthetic 10 000-point model vs 100 000-point scene, rate 1
  c3s a quarter-scale c3 (2 500 / 25 000) for quick runs
  c4  synthetic 8-model library (2 000 pts each) vs a 1 048 576-point scene, reference rate 20, one align per model
  c4s an eighth-scale c4 scene (131 072 pts) for quick runs

c1/c2 clouds are the reference's own data frozen under tests/golden (tools/make_fixtures.py: the
bottle PLY voxel-averaged, data/1_depth.exr back-projected with the reference's intrinsics because
data/1_cloud.ply is missing from the repository); c3 is seeded synthetic (synth.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import synth

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_GOLDEN = os.path.join(_ROOT, "tests", "golden")

ANGLE_STEP = np.float32(12.0) / np.float32(180.0) * np.float32(np.pi)  # PCL default 12 degrees
DIST_STEP = np.float32(0.01)
POS_THR = np.float32(0.01)
ROT_THR = np.float32(20.0 / 180.0 * np.pi)


@dataclass
class Workload:
    name: str
    description: str
    model: np.ndarray
    scene: np.ndarray
    ref_rate: int
    data: str
    angle_step: np.float32 = ANGLE_STEP
    dist_step: np.float32 = DIST_STEP
    pos_thr: np.float32 = POS_THR
    rot_thr: np.float32 = ROT_THR
    models: list = None  # model library (c4); None = the single `model`

    def library(self):
        return self.models if self.models else [self.model]

    @property
    def n_ref(self):
        return (self.scene.shape[0] + self.ref_rate - 1) // self.ref_rate


def _golden(name):
    return np.load(os.path.join(_GOLDEN, name + ".npz"))["cloud"].astype(np.float32)


_FIXTURE = "reference fixture (bottle_remesh_meter_normalized.ply voxel-averaged; scene back-projected from data/1_depth.exr)"


def load(name: str) -> Workload:
    if name == "c1":
        return Workload("c1", "bottle 1 cm (543 pts) vs YOLO-cropped 1_cloud (934 pts), PCL defaults, ref rate 5",
                        _golden("bottle_1cm"), _golden("scene_crop_1cm"), 5, _FIXTURE)
    if name == "c2":
        return Workload("c2", "bottle 1 cm (543 pts) vs full uncropped 1_cloud (44 893 pts), all reference points",
                        _golden("bottle_1cm"), _golden("scene_full_1cm"), 1, _FIXTURE)
    if name == "c2_5mm":
        return Workload("c2_5mm", "bottle 5 mm (2 009 pts) vs full uncropped 1_cloud (44 893 pts), all reference points",
                        _golden("bottle_5mm"), _golden("scene_full_1cm"), 1, _FIXTURE)
    if name == "c3":
        return Workload("c3", "synthetic 10 000-pt model vs 100 000-pt scene, all reference points",
                        synth.synth_model(10000, 1), synth.synth_scene(100000, 2, model_seed=1), 1, "synthetic")
    if name == "c3s":
        return Workload("c3s", "synthetic 2 500-pt model vs 25 000-pt scene, all reference points",
                        synth.synth_model(2500, 1), synth.synth_scene(25000, 2, model_seed=1), 1, "synthetic")
    if name in ("c4", "c4s"):
        n_s, rate = ((1 << 20), 20) if name == "c4" else ((1 << 17), 20)
        models = [synth.synth_model(2000, 10 + k, k) for k in range(8)]
        return Workload(name, f"synthetic 8-model library (2 000 pts each) vs {n_s}-pt scene (1024 x 1024 WFOV size), "
                              f"every 20th scene point a reference (the reference's 1/0.05), one align per model",
                        models[0], synth.synth_library_scene(n_s, 3), rate, "synthetic", models=models)
    raise ValueError(f"unknown workload {name!r}")
