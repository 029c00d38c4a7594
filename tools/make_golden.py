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
    # oracle_golden.npz was frozen in round 1, when the accumulator had floor(2*pi/step) columns with the votes of the
    # missing bin clamped into the last one (NALPHA_FLOOR_CLAMP); it stays the pin of that rule
    hm = ob.HashMap(A, D, nalpha_rule=ob.NALPHA_FLOOR_CLAMP).set_input_feature_cloud(feats)
    rng = np.random.default_rng(20261018)
    pairs = rng.choice(n * n, 512, replace=False)
    keys, lengths = hm.dump_keys()
    order = np.lexsort(keys.T[::-1])
    hyps, stats = hm.vote(model, scene, 0, 5, n_threads=1)
    poses, votes, assign, ncl = ob.cluster(hyps)
    acc, nv = hm.vote_accumulate(n, scene, 250)
    r1 = os.path.join(G, "oracle_golden.npz")
    np.savez_compressed(
        r1 if "--rewrite-round1" in sys.argv or not os.path.exists(r1) else os.path.join("/tmp", "oracle_golden_check.npz"),
        pair_index=pairs, pair_features=feats[pairs],
        n_entries=hm.num_entries, n_keys=hm.num_keys, model_diameter=np.float32(hm.model_diameter),
        keys=keys[order], key_lengths=lengths[order],
        hyp_votes=hyps["votes"], hyp_model_index=hyps["model_index"], hyp_alpha_bin=hyps["alpha_bin"],
        hyp_pose=hyps["pose"], vote_stats=np.array([stats[k] for k in ("pairs_examined", "pairs_in_radius", "nonempty_lookups", "votes")]),
        cluster_poses=poses, cluster_votes=votes, cluster_assign=assign, n_clusters=ncl,
        acc250_nonzero_index=np.flatnonzero(acc.reshape(-1)), acc250_nonzero_value=acc.reshape(-1)[acc.reshape(-1) > 0],
    )
    print("entries", hm.num_entries, "keys", hm.num_keys, "votes", stats, "clusters", ncl, votes)
    # the other two column rules (ceil = the default since round 2; floor with the votes dropped = PCL <= 1.11)
    out = {}
    for name, rule in (("ceil", ob.NALPHA_CEIL), ("drop", ob.NALPHA_FLOOR_DROP)):
        hm.set_nalpha_rule(rule)
        hyps, stats = hm.vote(model, scene, 0, 5, n_threads=1)
        poses, votes, assign, ncl = ob.cluster(hyps)
        acc, nv = hm.vote_accumulate(n, scene, 250)
        out.update({f"{name}_hyp_votes": hyps["votes"], f"{name}_hyp_model_index": hyps["model_index"],
                    f"{name}_hyp_alpha_bin": hyps["alpha_bin"], f"{name}_hyp_pose": hyps["pose"],
                    f"{name}_votes_cast": np.uint64(stats["votes"]), f"{name}_cluster_votes": votes,
                    f"{name}_cluster_poses": poses, f"{name}_n_clusters": ncl,
                    f"{name}_acc250_nonzero_index": np.flatnonzero(acc.reshape(-1)),
                    f"{name}_acc250_nonzero_value": acc.reshape(-1)[acc.reshape(-1) > 0], f"{name}_n_alpha": acc.shape[1]})
        print(name, "columns", acc.shape[1], "clusters", ncl, votes)
    np.savez_compressed(os.path.join(G, "oracle_golden_rules.npz"), **out)


if __name__ == "__main__":
    main()
