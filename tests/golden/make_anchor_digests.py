#!/usr/bin/env python
"""Mint tests/golden/anchor_digests.json (AUTHORING container only): SHA-256 of the float32 anchors the reference's
FpnAnchorGenerator.generate_anchors (fpn_anchor_generator.py:21-59, levels 3..7 concatenated as
bdd_dataset_handler.py:161-186 does) produces for the image shapes of BASELINE.json's configs and a few odd ones.

    python tests/golden/make_anchor_digests.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import tf_numpy_shim as shim  # noqa: E402

SHAPES = [(720, 1280), (375, 1242), (512, 1696), (370, 1224), (97, 161), (64, 96), (33, 47)]


def main():
    _, _, ag, _, _ = shim.load_reference()
    gen = ag.FpnAnchorGenerator(dict(aspect_ratios=[[1.0, 1.0], [1.0, 2.0], [2.0, 1.0]], scales=[1.0, 1.26, 1.59]))
    out = {}
    for h, w in SHAPES:
        ref = np.concatenate([np.asarray(gen.generate_anchors(shim._t(np.asarray((h, w, 3), np.int32)), l))
                              for l in [3, 4, 5, 6, 7]], axis=0).astype(np.float32)
        out[f"{h}x{w}"] = {"A": int(ref.shape[0]), "sha256": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()}
        print(h, w, out[f"{h}x{w}"])
    json.dump(out, open(os.path.join(HERE, "anchor_digests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
