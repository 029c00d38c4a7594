#!/usr/bin/env python
"""Freezes oracle outputs on the committed fixtures into tests/golden/oracle_golden.npz.

PARITY UNPINNED: the reference repository ships no golden vectors for this path and
PCL cannot be built here, so these vectors pin the ORACLE against regressions (and record the key
statistics SURVEY.md Appendix C obtained from an independent numpy restatement), nothing more.
Run in the build container:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
A = np.float32(12.0) / np.float32(180.0) * np.float32(np.pi)
D = np.float32(0.01)


def main():
    model = np.load(os.path.join(G, "bottle_1cm.npz"))["cloud"]
    scene = np.load(os.path.join(G, "scene_crop_1cm.npz"))["cloud"]
    n = model.shape[0]
    feats = ob.ppf_estimation(model)
    hm = ob.HashMap(A, D).set_input_feature_cloud(feats)
    rng = np.random.default_rng(20261018)
    pairs = rng.choice(n * n, 512, replace=False)
    keys, lengths = hm.dump_keys()
    order = np.lexsort(keys.T[::-1])
    hyps, stats = hm.vote(model, scene, 0, 5, n_threads=1)
    poses, votes, assign, ncl = ob.cluster(hyps)
    acc, nv = hm.vote_accumulate(n, scene, 250)
    np.savez_compressed(
        os.path.join(G, "oracle_golden.npz"),
        pair_index=pairs, pair_features=feats[pairs],
        n_entries=hm.num_entries, n_keys=hm.num_keys, model_diameter=np.float32(hm.model_diameter),
        keys=keys[order], key_lengths=lengths[order],
        hyp_votes=hyps["votes"], hyp_model_index=hyps["model_index"], hyp_alpha_bin=hyps["alpha_bin"],
        hyp_pose=hyps["pose"], vote_stats=np.array([stats[k] for k in ("pairs_examined", "pairs_in_radius", "nonempty_lookups", "votes")]),
        cluster_poses=poses, cluster_votes=votes, cluster_assign=assign, n_clusters=ncl,
        acc250_nonzero_index=np.flatnonzero(acc.reshape(-1)), acc250_nonzero_value=acc.reshape(-1)[acc.reshape(-1) > 0],
    )
    print("entries", hm.num_entries, "keys", hm.num_keys, "votes", stats, "clusters", ncl, votes)


if __name__ == "__main__":
    main()
