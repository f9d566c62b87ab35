#!/usr/bin/env python
"""Mint tests/golden/writers_{bdd,kitti}.npz (AUTHORING container only, needs /root/reference).

Executes the reference's own formatting functions verbatim over the numpy `tf` shim

    validation_utils.predictions_to_bdd_format     # validation_utils.py:183-213
    validation_utils.predictions_to_kitti_format   # validation_utils.py:216-272
    box_utils.vuhw_to_vuvu_np                      # box_utils.py:70-88

and writes their results exactly as run_inference.py does: json.dump(final_results_list, fp,
indent=4, separators=(',', ': ')) (:258-260) and np.savetxt(name, rows, newline='\\r\\n', fmt='%s')
/ np.savetxt(name, []) (:195-201).  The fixtures hold the padded result blocks and the bytes of
every file.

    python tests/golden/make_writer_golden.py
"""
import importlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import tf_numpy_shim as shim  # noqa: E402

BDD_CATEGORIES = ['car', 'bus', 'truck', 'person', 'rider', 'bike', 'motor']      # bdd_dataset_handler.py categories (K-1 = 7)


def make_block(seed, B, D, K, num_dets):
    rng = np.random.default_rng(seed)
    means = np.stack([rng.uniform(0, 700, (B, D)), rng.uniform(0, 1200, (B, D)),
                      np.exp(rng.uniform(0, 6, (B, D))), np.exp(rng.uniform(0, 6, (B, D)))], axis=2).astype(np.float32)
    # number-format corner cases: integers, tiny / huge magnitudes, negatives, exact halves, zero
    special = np.array([[100., 200., 50., 80.], [1e-5, 2.5e-7, 3e-5, 1e-4], [1e16, 3e17, 2e16, 5e15],
                        [-3.25, -1e-3, 7., 9.], [0., 0., 0., 0.], [123456.79, 9999999., 16777216., 0.1],
                        [1e-4, 9.9e-5, 2e-4, 1.0001e-4]], np.float32)
    n = min(len(special), D)
    means[0, :n] = special[:n]
    cat = rng.dirichlet(np.ones(K) * 0.3, (B, D)).astype(np.float32)
    cat[0, 0] = 0.125                      # all equal: first maximum is class 0
    if D > 3:
        cat[0, 1, :] = 0; cat[0, 1, K - 1] = 1.0          # background wins: skipped by both formats
        cat[0, 2, :] = 0; cat[0, 2, 1] = 0.5; cat[0, 2, 2] = 0.5   # tie between 1 and 2 -> 1
        cat[0, 3, :] = 1e-8; cat[0, 3, 0] = 3e-5
    return np.asarray(num_dets, np.int32), means, cat


def main():
    shim.load_reference()
    vu = importlib.import_module("src.retina_net.experiments.validation_utils")
    bu = importlib.import_module("src.retina_net.anchor_generator.box_utils")

    # ---- bdd: one json over two result blocks (appended in order), K = 8 ----
    blocks = [make_block(1, 3, 12, 8, [12, 0, 5]), make_block(2, 2, 12, 8, [1, 9])]
    ids = [["b1c66a42-6f7d68ca.jpg", "weird \"name\"\\é中\U0001F600.jpg", "c.jpg"], ["d\t.jpg", "e.jpg"]]
    final_results_list = []
    for (nd, means, cat), names in zip(blocks, ids):
        for b in range(len(nd)):
            boxes = bu.vuhw_to_vuvu_np(means[b, :nd[b]])
            final_results_list.extend(vu.predictions_to_bdd_format(boxes, cat[b, :nd[b]], names[b], category_list=BDD_CATEGORIES))
    fp = io.StringIO()
    json.dump(final_results_list, fp, indent=4, separators=(',', ': '))
    out = {"json": np.frombuffer(fp.getvalue().encode(), np.uint8), "categories": np.array(BDD_CATEGORIES),
           "n_blocks": np.int32(len(blocks))}
    for i, ((nd, means, cat), names) in enumerate(zip(blocks, ids)):
        out.update({f"num_dets{i}": nd, f"means{i}": means, f"cat_param{i}": cat, f"ids{i}": np.array(names)})
    fp = io.StringIO()
    json.dump([], fp, indent=4, separators=(',', ': '))
    out["json_empty"] = np.frombuffer(fp.getvalue().encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "writers_bdd.npz"), **out)
    print("writers_bdd.npz", len(final_results_list), "entries,", len(out["json"]), "bytes")

    # ---- kitti: one txt per image, K = 4 (car, pedestrian, cyclist, background) ----
    nd, means, cat = make_block(3, 4, 10, 4, [10, 0, 3, 1])
    cat[3, 0] = [0.1, 0.2, 0.6, 0.1]      # the only detection is a cyclist: rows empty -> np.savetxt(name, [])
    names = ["000001", "000002", "000003", "000004"]
    out = {"num_dets": nd, "means": means, "cat_param": cat, "ids": np.array(names)}
    for b in range(len(nd)):
        boxes = bu.vuhw_to_vuvu_np(means[b, :nd[b]])
        rows = vu.predictions_to_kitti_format(boxes, cat[b, :nd[b]])
        fp = io.BytesIO()
        if rows.size == 0:
            np.savetxt(fp, [])
        else:
            np.savetxt(fp, rows, newline='\r\n', fmt='%s')
        out[f"txt{b}"] = np.frombuffer(fp.getvalue(), np.uint8)
        print(names[b], rows.shape, len(fp.getvalue()), "bytes")
    np.savez_compressed(os.path.join(HERE, "writers_kitti.npz"), **out)


if __name__ == "__main__":
    main()
