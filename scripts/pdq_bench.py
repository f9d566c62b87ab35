"""PDQ spatial quality on the GPU vs the CPU oracle (reference algorithm, one core): synthetic BDD-shape
evaluation load — 720x1280 images, ~60 kept detections and ~20 ground-truth boxes per image, corner
covariances of a few to a few hundred px^2 (fused covariances x70, compute_pdq.py:98-103).
Prints one JSON object.  Usage: python scripts/pdq_bench.py [n_images] [dets] [gts]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayes_od_rc_b200 import pdq as ppdq  # noqa: E402

H, W = 720, 1280


def scene(rng, D, G):
    boxes, covs = [], []
    for _ in range(D):
        h, w = np.exp(rng.uniform(np.log(24), np.log(400))), np.exp(rng.uniform(np.log(24), np.log(400)))
        h, w = min(h, H - 2), min(w, W - 2)
        y1, x1 = rng.uniform(0, H - h - 1), rng.uniform(0, W - w - 1)
        boxes.append([x1, y1, x1 + w, y1 + h])
        cs = []
        for _ in range(2):
            a = rng.normal(size=(2, 2))
            cs.append((a @ a.T + 0.5 * np.eye(2)) * np.exp(rng.uniform(np.log(2), np.log(150))))
        covs.append(cs)
    gt = []
    for _ in range(G):
        h, w = np.exp(rng.uniform(np.log(24), np.log(400))), np.exp(rng.uniform(np.log(24), np.log(400)))
        h, w = min(h, H - 2), min(w, W - 2)
        y1, x1 = rng.uniform(0, H - h), rng.uniform(0, W - w)
        gt.append([x1, y1, x1 + w, y1 + h])
    return np.array(boxes).astype(np.int32), np.array(covs), np.array(gt).astype(np.int32)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_img = int(args[0]) if len(args) > 0 else 64
    D = int(args[1]) if len(args) > 1 else 60
    G = int(args[2]) if len(args) > 2 else 20
    reps = int(os.environ.get('PDQ_REPS', '10'))
    rng = np.random.default_rng(0)
    scenes = [scene(rng, D, G) for _ in range(n_img)]
    do = np.arange(n_img + 1) * D
    go = np.arange(n_img + 1) * G
    boxes = np.concatenate([s[0] for s in scenes]); covs = np.concatenate([s[1] for s in scenes]); gts = np.concatenate([s[2] for s in scenes])
    eng = ppdq.PdqEngine((H, W))
    for _ in range(min(3, reps)):
        eng.losses(do, boxes, covs, go, gts)
    t = time.perf_counter()
    dev = []
    for _ in range(reps):
        eng.losses(do, boxes, covs, go, gts)
        dev.append(eng.last_ms())
    wall = (time.perf_counter() - t) / reps
    ms = {k: float(np.median([d[k] for d in dev])) for k in ("roi", "tables", "sums_or_maps")}
    out = {"workload": f"{n_img} images 720x1280, {D} detections + {G} ground-truth boxes each",
           "losses": {"wall_ms_per_call": wall * 1e3, "images_per_s": n_img / wall, "device_ms": ms,
                      "table_floats": dev[-1]["table_floats"], "launches": dev[-1]["launches"]}}
    # dense maps into device memory: HBM-write bound
    import torch
    nd = min(len(boxes), 256)
    buf = torch.empty((nd, H, W), device="cuda")
    for _ in range(min(3, reps)):
        eng.heatmaps(boxes[:nd], covs[:nd], out=buf)
    dm = []
    for _ in range(reps):
        eng.heatmaps(boxes[:nd], covs[:nd], out=buf)
        dm.append(eng.last_ms()["sums_or_maps"])
    m = float(np.median(dm))
    out["dense_maps"] = {"detections": nd, "kernel_ms": m, "bytes_written": nd * H * W * 4, "GBps": nd * H * W * 4 / m / 1e6}
    # CPU: the oracle (reference algorithm restated in C, one core) on a bounded sample
    if "--no-cpu" not in sys.argv:
        from oracle import pdq as opdq
        ns = 2
        t = time.perf_counter()
        for i in range(ns):
            b, c, g = scenes[i]
            hm = opdq.heatmaps((H, W), b, c)
            fg, bg, tot = opdq.losses(hm, g)
        cpu = (time.perf_counter() - t) / ns
        f2, b2, t2 = eng.losses([0, D], b, c, [0, G], g)
        out["cpu_oracle"] = {"s_per_image": cpu, "images_per_s": 1 / cpu, "cores": 1, "sample": f"{ns} images",
                             "max_rel_diff_fg": float(np.max(np.abs(f2[0] - fg) / np.maximum(np.abs(fg), 1e-9))),
                             "max_rel_diff_bg": float(np.max(np.abs(b2[0] - bg) / np.maximum(np.abs(bg), 1e-9)))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
