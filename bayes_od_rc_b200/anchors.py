"""Host-side FPN anchor grid, mirroring the reference's producer of
``sample_dict['anchors']``.

Follows ``FpnAnchorGenerator.generate_anchors``
(src/retina_net/anchor_generator/fpn_anchor_generator.py:21-59) and the P3->P7
concatenation of ``BddDatasetHandler.create_sample_dict``
(src/retina_net/datasets/bdd/bdd_dataset_handler.py:161-186): level l has stride
2^l and base side 2^(l+2); positions are ``(arange(ceil(dim/stride)) + 0.5) *
stride`` with u fastest; 9 anchors per location ordered aspect-major,
scale-minor; every product is taken in float32 in the reference's order.
The device-side twin is ``bod_generate_anchors`` (include/bayesod.h).
"""
from __future__ import annotations

import numpy as np

ASPECT_RATIOS = ((1.0, 1.0), (1.0, 2.0), (2.0, 1.0))   # retinanet_bdd.yaml:56
SCALES = (1.0, 1.26, 1.59)                               # retinanet_bdd.yaml:58
LEVELS = (3, 4, 5, 6, 7)                                 # retinanet_bdd.yaml:53


def anchor_dims(level: int, aspect_ratios=ASPECT_RATIOS, scales=SCALES) -> np.ndarray:
    """[9,2] (h,w) of the anchors of one pyramid level (fpn_anchor_generator.py:35-48)."""
    f32 = np.float32
    side = f32(2.0 ** (level + 2))
    dims = []
    for rh, rw in aspect_ratios:
        for s in scales:
            if rh == 1 and rw == 1:
                dims.append([f32(rh) * side * f32(s), f32(rw) * side * f32(s)])
            else:
                sol = np.sqrt((side * side) / f32(rh * rw), dtype=f32)
                dims.append([f32(rh) * sol * f32(s), f32(rw) * sol * f32(s)])
    return np.asarray(dims, dtype=f32)


def level_grid(im_h: int, im_w: int, level: int):
    stride = 2 ** level
    return -(-im_h // stride), -(-im_w // stride)       # ceil, as tf.range(0, dim/stride) yields


def generate_anchors(im_h: int, im_w: int, levels=LEVELS) -> np.ndarray:
    """[A,4] float32 anchors (v,u,h,w), P3->P7."""
    out = []
    for level in levels:
        stride = np.float32(2 ** level)
        nv, nu = level_grid(im_h, im_w, level)
        v = (np.arange(nv, dtype=np.float32) + np.float32(0.5)) * stride
        u = (np.arange(nu, dtype=np.float32) + np.float32(0.5)) * stride
        uu, vv = np.meshgrid(u, v)
        loc = np.stack([vv.reshape(-1), uu.reshape(-1)], axis=1)         # [L,2]
        dims = anchor_dims(level)
        grid = np.concatenate([np.repeat(loc, len(dims), axis=0), np.tile(dims, (len(loc), 1))], axis=1)
        out.append(grid.astype(np.float32))
    return np.concatenate(out, axis=0)


def num_anchors(im_h: int, im_w: int, levels=LEVELS) -> int:
    return sum(9 * nv * nu for nv, nu in (level_grid(im_h, im_w, l) for l in levels))


def level_anchor_counts(im_h: int, im_w: int, levels=LEVELS) -> list:
    """Anchors per pyramid level, P3->P7 (the ``level_anchors`` of ``bod_config`` for ``bod_run_levels``)."""
    return [9 * nv * nu for nv, nu in (level_grid(im_h, im_w, l) for l in levels)]
