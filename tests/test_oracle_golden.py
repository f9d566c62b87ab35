"""CPU: pin the C oracle against the golden fixtures, i.e. against the
reference's own source files executed over the numpy TF shim
(tests/golden/make_golden.py).  Index outputs must be bit-exact; floating
outputs within rtol 1e-4 / atol 1e-5 (north_star), with binary64 adjudication
for ill-conditioned elements."""
import numpy as np
import pytest

import oracle
from helpers import (adjudicated_close, check_categorical_merge, golden_cases, load_golden, oracle_config_of,
                     val_golden_cases, val_scaling_of, within_tol)

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("name", CASES)
def test_anchors_bit_exact(name):
    g = load_golden(name)
    h, w = g["meta"]["image_shape"]
    mine = oracle.generate_anchors(h, w)
    assert mine.shape == g["anchors"].shape
    assert np.array_equal(mine.view(np.uint32), g["anchors"].view(np.uint32))


def test_anchors_full_size_digests():
    """Oracle and the host mirror (bayes_od_rc_b200/anchors.py) against SHA-256 digests of the reference generator's output
    for the image shapes of BASELINE.json's configs and some odd ones (tests/golden/make_anchor_digests.py)."""
    import hashlib
    import json
    import os
    from helpers import GOLDEN_DIR
    from bayes_od_rc_b200 import anchors as host_anchors
    digests = json.load(open(os.path.join(GOLDEN_DIR, "anchor_digests.json")))
    assert len(digests) >= 7
    for key, want in digests.items():
        h, w = (int(v) for v in key.split("x"))
        for name, arr in (("oracle", oracle.generate_anchors(h, w)), ("host mirror", host_anchors.generate_anchors(h, w))):
            arr = np.ascontiguousarray(arr, np.float32)
            assert arr.shape == (want["A"], 4), (key, name)
            assert hashlib.sha256(arr.tobytes()).hexdigest() == want["sha256"], (key, name)
        assert host_anchors.num_anchors(h, w) == want["A"]


@pytest.mark.parametrize("name", CASES)
def test_inference_half(name):
    """bayes_od_inference outputs (inference_utils.py:217)."""
    g = load_golden(name)
    cfg = oracle_config_of(g["meta"])
    r = oracle.run_image(cfg, g["cls"], g["box"], g["cov"] if g["meta"]["has_cov"] else None, g["anchors"], g["counts"])
    S = len(g["cnt_post"])
    assert len(r.keep) == S
    if S == 0:
        assert len(r.nms_indices) == 0
        return
    # counts: small integers + 1/K -> exact
    assert np.array_equal(r.cnt_post, g["cnt_post"])
    mu_ref = g["mu_post"][:, :, 0]
    assert within_tol(r.mu_post, mu_ref).all(), np.abs(r.mu_post - mu_ref).max()
    r64 = oracle.run_image(cfg, g["cls"], g["box"], g["cov"] if g["meta"]["has_cov"] else None, g["anchors"], g["counts"],
                           real="f64", force=dict(nms_indices=r.nms_indices, mask=r.mask))
    ok, frac = adjudicated_close(r.sig_post, g["sig_post"], r64.sig_post)
    assert ok.all(), (frac, np.abs(r.sig_post - g["sig_post"]).max())
    assert frac > 0.99
    # centre selection: bit-exact, in order
    assert np.array_equal(r.nms_indices, g["nms_indices"]), (r.nms_indices[:20], g["nms_indices"][:20])
    # membership of every cluster: bit-exact
    mem = oracle.mask_to_bool(r.mask, S)                     # [D,S]
    ref_mem = (g["iou_cols"] > cfg.iou_threshold).T            # iou[:, centre] > thr (:316)
    assert np.array_equal(mem, ref_mem)


@pytest.mark.parametrize("name", CASES)
def test_clustering_half(name):
    """bayes_od_clustering outputs (inference_utils.py:364)."""
    g = load_golden(name)
    if "final_means" not in g:
        pytest.skip("no survivors in this fixture")
    cfg = oracle_config_of(g["meta"])
    cov = g["cov"] if g["meta"]["has_cov"] else None
    r = oracle.run_image(cfg, g["cls"], g["box"], cov, g["anchors"], g["counts"])
    r64 = oracle.run_image(cfg, g["cls"], g["box"], cov, g["anchors"], g["counts"], real="f64",
                           force=dict(nms_indices=r.nms_indices, mask=r.mask))
    assert r.empty_clusters == 0
    assert within_tol(r.final_means, g["final_means"][:, :, 0]).all()
    ok, frac = adjudicated_close(r.final_covs, g["final_covs"], r64.final_covs)
    assert ok.all() and frac > 0.99, frac
    mem = oracle.mask_to_bool(r.mask, len(r.keep))
    n_plain, n_tied = check_categorical_merge(g, r.nms_indices, mem, r.extra["chosen"], r.final_scores, r.final_counts)
    assert n_plain > 0


@pytest.mark.parametrize("name", CASES)
def test_stage_isolated_nms_and_clustering(name):
    """Feed the GOLDEN posterior (means, covs, counts) to the oracle's NMS +
    clustering: removes upstream float noise, so everything index-like is exact."""
    g = load_golden(name)
    S = len(g["cnt_post"])
    if S == 0:
        pytest.skip("no survivors")
    cfg = oracle_config_of(g["meta"])
    mu = g["mu_post"][:, :, 0]
    corners = np.stack([mu[:, 0] - mu[:, 2] / np.float32(2), mu[:, 1] - mu[:, 3] / np.float32(2),
                        mu[:, 0] + mu[:, 2] / np.float32(2), mu[:, 1] + mu[:, 3] / np.float32(2)], 1).astype(np.float32)
    if cfg.ranking_method == "score":
        tot = g["cnt_post"][:, 0].copy()                      # row sums in index order (numpy's pairwise
        for k in range(1, g["cnt_post"].shape[1]):            # order differs by an ulp when 1/K is inexact, e.g. K = 11)
            tot = tot + g["cnt_post"][:, k]
        p = g["cnt_post"] / tot[:, None]
        score = p.max(1).astype(np.float32)
        sel, _ = oracle.nms_v5(corners, score, cfg.max_output_size, cfg.iou_threshold, -np.inf, cfg.soft_nms_sigma)
        assert np.array_equal(sel, g["nms_indices"])
    sel = g["nms_indices"]
    mask = oracle.membership(corners, sel, cfg.iou_threshold)
    assert np.array_equal(oracle.mask_to_bool(mask, S), (g["iou_cols"] > cfg.iou_threshold).T)
    fs, fm, fc, fn, mem, empty = oracle.clustering(g["cnt_post"], mu, g["sig_post"], sel, mask, 70.0)
    assert empty == 0
    assert within_tol(fm, g["final_means"][:, :, 0]).all()
    check_categorical_merge(g, sel, oracle.mask_to_bool(mask, S), oracle.clustering.last_chosen, fs, fn)
    fs64, fm64, fc64, *_ = oracle.clustering(g["cnt_post"], mu, g["sig_post"], sel, mask, 70.0, real="f64")
    ok, frac = adjudicated_close(fc, g["final_covs"], fc64)
    assert ok.all() and frac > 0.99


@pytest.mark.parametrize("name", val_golden_cases())
def test_validation_postprocess(name):
    """validation_utils.post_process_predictions (validation_utils.py:10-77), executed verbatim over the shim,
    against the oracle restatement: same boxes selected in the same order, values within tolerance."""
    g = load_golden(name)
    mode, shift, norm_hw, scale_hw = val_scaling_of(g["meta"])
    r = oracle.val_postprocess(g["cls"], g["box"], g["anchors"], scale_mode=mode, shift=shift, norm_hw=norm_hw, scale_hw=scale_hw)
    D = len(g["classes_out"])
    assert len(r.nms_indices) == D
    if D == 0:
        assert len(r.keep) == 0
        return
    assert within_tol(r.classes_out, g["classes_out"]).all()
    assert within_tol(r.corners_out, g["corners_out"], rtol=1e-5, atol=1e-4).all()
    # kept anchors = softmax arg-max is not the background column (float-independent here: margins are large)
    keep = np.flatnonzero(np.argmax(g["cls"], axis=1) != g["cls"].shape[1] - 1)
    assert np.array_equal(r.keep, keep)


def test_validation_fixtures_present():
    assert len(val_golden_cases()) >= 3
