#!/usr/bin/env python
"""Shared-memory reduction rates on this GPU (csrc/microbench.cu): the denominators of the K3 roofline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_ppf_pose_estimation_b200 import capi
names = {0: "conflict_free", 1: "random_lcg", 2: "one_word", 3: "two_per_bank", 4: "four_per_bank",
         5: "gather_conflict_free", 6: "gather_random", 7: "gather_two_per_bank"}
ctx = capi.Context(0)
out = {}
for shape, off in (("512x2", 0), ("1024x1", 8)):
    for p, n in names.items():
        out[f"{n}_{shape}"] = max(ctx.microbench_atoms(p + off) for _ in range(3))
print(json.dumps(out))
