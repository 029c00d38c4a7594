"""Shared fixtures.  Tests that need a B200 carry @pytest.mark.gpu; everything else runs on CPU.

The oracle (oracle/) is test infrastructure: it is imported here and in the tests only.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

ANGLE_STEP = np.float32(12.0) / np.float32(180.0) * np.float32(np.pi)  # PCL default, float arithmetic
DIST_STEP = np.float32(0.01)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_cloud(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))["cloud"].astype(np.float32)


@pytest.fixture(scope="session")
def bottle():
    return load_cloud("bottle_1cm")


@pytest.fixture(scope="session")
def bottle_5mm():
    return load_cloud("bottle_5mm")


@pytest.fixture(scope="session")
def scene_crop():
    return load_cloud("scene_crop_1cm")


@pytest.fixture(scope="session")
def scene_full():
    return load_cloud("scene_full_1cm")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def oracle_bottle(oracle, bottle):
    """(features, hashmap) of the 1 cm bottle, PCL defaults."""
    feats = oracle.ppf_estimation(bottle)
    hm = oracle.HashMap(ANGLE_STEP, DIST_STEP).set_input_feature_cloud(feats)
    return feats, hm


@pytest.fixture(scope="session")
def ctx():
    """A device context; only reachable from gpu-marked tests."""
    from yolo_ppf_pose_estimation_b200 import capi
    return capi.Context(0)
