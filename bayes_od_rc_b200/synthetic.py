"""Deterministic synthetic RetinaNet head outputs.

There is no dataset or checkpoint in this environment, so the bench and the
parity tests feed the path with head tensors shaped like what
``RetinaNetModel.call(..., 'testing')`` produces
(src/retina_net/models/retinanet_model.py:73-112):

    cls  [N,A,K]   logits, background = last column, bias init -log(99) on the
                   foreground columns (src/retina_net/models/multitask_headers.py:79-83)
    box  [N,A,4]   regression deltas in the encoding of
                   fpn_anchor_generator.py:103-117 (x10 on centres, x5 on log sizes)
    cov  [N,A,4,4] lower-triangular, log-variance on the diagonal
                   (or the packed [N,A,10] vector before tfp.math.fill_triangular)

A scene of G ground-truth boxes is drawn per image; anchors whose IoU with a
ground-truth box is >= fg_iou fire on that class and regress onto it, the rest
look like background, and a small fraction of "stray" background anchors fire
on a random class.  Each of the N MC-dropout samples adds i.i.d. noise.
Everything is generated with a per-image ``torch.Generator`` seeded
``1000 * config_id + image_id`` so any image can be regenerated anywhere.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import anchors as anchors_mod


@dataclass
class SceneSpec:
    im_h: int = 720
    im_w: int = 1280
    N: int = 10                 # MC-dropout samples
    K: int = 8                  # classes + background
    g_min: int = 30
    g_max: int = 60
    fg_iou: float = 0.4
    fg_logit: float = 4.0
    bg_logit_for_fg: float = -2.0
    stray_frac: float = 0.002
    box_lo: float = 24.0
    box_hi: float = 400.0
    config_id: int = 3
    packed_cov: bool = False    # emit [N,A,10] instead of [N,A,4,4]


# tfp.math.fill_triangular index map for n=4 (retinanet_model.py:110):
# packed x0..x9 -> L[i][j]; PACKED_OF[i][j] is the packed index feeding L[i][j].
PACKED_OF = {(0, 0): 4, (1, 0): 8, (1, 1): 9, (2, 0): 7, (2, 1): 6, (2, 2): 5,
             (3, 0): 3, (3, 1): 2, (3, 2): 1, (3, 3): 0}


def fill_triangular(packed: torch.Tensor) -> torch.Tensor:
    """[...,10] -> [...,4,4] lower triangular, TFP order."""
    out = packed.new_zeros(packed.shape[:-1] + (4, 4))
    for (i, j), p in PACKED_OF.items():
        out[..., i, j] = packed[..., p]
    return out


def _pair_iou_vuhw(a: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """Ordinary IoU between anchors [A,4] and boxes [G,4], both (v,u,h,w)."""
    a0 = a[:, None, 0] - a[:, None, 2] / 2; a2 = a[:, None, 0] + a[:, None, 2] / 2
    a1 = a[:, None, 1] - a[:, None, 3] / 2; a3 = a[:, None, 1] + a[:, None, 3] / 2
    g0 = g[None, :, 0] - g[None, :, 2] / 2; g2 = g[None, :, 0] + g[None, :, 2] / 2
    g1 = g[None, :, 1] - g[None, :, 3] / 2; g3 = g[None, :, 1] + g[None, :, 3] / 2
    ih = (torch.minimum(a2, g2) - torch.maximum(a0, g0)).clamp_min(0)
    iw = (torch.minimum(a3, g3) - torch.maximum(a1, g1)).clamp_min(0)
    inter = ih * iw
    return inter / (a[:, None, 2] * a[:, None, 3] + g[None, :, 2] * g[None, :, 3] - inter)


def make_image(spec: SceneSpec, image_id: int, anchors: torch.Tensor, device="cpu",
               with_counts: bool = True, num_draws: int = 30):
    """One image worth of head outputs on `device`.

    Returns dict(cls [N,A,K], box [N,A,4], cov [N,A,4,4] | [N,A,10], counts [A,K] | None).
    Random numbers are always drawn on the CPU generator stream of the image when
    device == 'cpu'; on CUDA a CUDA generator with the same seed is used (the two
    streams differ — parity tests pass ONE set of tensors to both sides)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 * spec.config_id + image_id)
    A = anchors.shape[0]
    N, K = spec.N, spec.K
    C_fg = K - 1
    anc = anchors.to(dev)

    def rnd(*shape):
        return torch.randn(*shape, generator=gen, device=dev, dtype=torch.float32)

    def uni(*shape):
        return torch.rand(*shape, generator=gen, device=dev, dtype=torch.float32)

    # 1. scene
    G = int(torch.randint(spec.g_min, spec.g_max + 1, (1,), generator=gen, device=dev).item())
    gv = uni(G) * spec.im_h
    gu = uni(G) * spec.im_w
    lo, hi = math.log(spec.box_lo), math.log(spec.box_hi)
    gh = torch.exp(lo + uni(G) * (hi - lo))
    gw = torch.exp(lo + uni(G) * (hi - lo))
    gcls = torch.randint(0, C_fg, (G,), generator=gen, device=dev)
    gt = torch.stack([gv, gu, gh, gw], dim=1)

    # 3. per-anchor base
    iou = _pair_iou_vuhw(anc, gt)
    best_iou, best = iou.max(dim=1)
    fg = best_iou >= spec.fg_iou
    tgt = gt[best]
    base_cls = torch.empty(A, K, device=dev)
    base_cls[:, :C_fg] = -math.log(99.0)
    base_cls[:, C_fg] = 0.0
    base_cls += 0.5 * rnd(A, K)
    fg_rows = torch.zeros(A, K, device=dev)
    fg_rows[:, C_fg] = spec.bg_logit_for_fg
    fg_rows.scatter_(1, gcls[best][:, None], spec.fg_logit)
    base_cls = torch.where(fg[:, None], fg_rows, base_cls)
    # strays
    stray = (~fg) & (uni(A) < spec.stray_frac)
    stray_cls = torch.randint(0, C_fg, (A,), generator=gen, device=dev)
    bump = torch.zeros(A, K, device=dev)
    bump.scatter_(1, stray_cls[:, None], math.log(99.0) + 1.0)
    base_cls = base_cls + bump * stray[:, None]

    # regression targets, fpn_anchor_generator.py:103-117
    t_exact = torch.stack([(tgt[:, 0] - anc[:, 0]) / anc[:, 2] * 10.0,
                           (tgt[:, 1] - anc[:, 1]) / anc[:, 3] * 10.0,
                           torch.log(tgt[:, 2] / anc[:, 2]) * 5.0,
                           torch.log(tgt[:, 3] / anc[:, 3]) * 5.0], dim=1)
    base_box = torch.where(fg[:, None], t_exact + 0.3 * rnd(A, 4), rnd(A, 4))

    # 4. MC spread
    cls = base_cls[None] + 0.25 * rnd(N, A, K)
    box = base_box[None] + 0.15 * rnd(N, A, 4)

    # 5. covariance head (packed 10-vector; diagonal slots get the log-variance)
    diag_slots = [PACKED_OF[(i, i)] for i in range(4)]
    base_cov = 0.1 * rnd(A, 10)
    base_cov[:, diag_slots] = 1.5 + 0.5 * rnd(A, 4)
    cov_packed = base_cov[None] + 0.05 * rnd(N, A, 10)
    cov = cov_packed if spec.packed_cov else fill_triangular(cov_packed)

    out = dict(cls=cls.contiguous(), box=box.contiguous(), cov=cov.contiguous(), counts=None,
               num_gt=G, num_fg=int(fg.sum().item()), num_stray=int(stray.sum().item()))
    # 6. counts for parity runs
    if with_counts:
        probs = torch.softmax(cls, dim=2).mean(dim=0)
        draws = torch.multinomial(probs, num_draws, replacement=True, generator=gen)
        counts = torch.zeros(A, K, device=dev)
        counts.scatter_add_(1, draws, torch.ones_like(draws, dtype=torch.float32))
        out["counts"] = counts
    return out


def make_batch(spec: SceneSpec, B: int, device="cpu", with_counts=True, first_image_id=0, num_draws=30):
    """Stack B images: cls [B,N,A,K], box [B,N,A,4], cov [B,N,A,4,4|10], counts [B,A,K], anchors [A,4]."""
    anc = torch.from_numpy(anchors_mod.generate_anchors(spec.im_h, spec.im_w))
    imgs = [make_image(spec, first_image_id + b, anc, device, with_counts, num_draws) for b in range(B)]
    out = dict(anchors=anc.to(device),
               cls=torch.stack([i["cls"] for i in imgs]),
               box=torch.stack([i["box"] for i in imgs]),
               cov=torch.stack([i["cov"] for i in imgs]),
               counts=torch.stack([i["counts"] for i in imgs]) if with_counts else None,
               meta=[{k: i[k] for k in ("num_gt", "num_fg", "num_stray")} for i in imgs])
    return out


def to_numpy(batch: dict) -> dict:
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
