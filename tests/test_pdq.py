"""PDQ spatial quality: oracle/pdq_oracle.c and the CUDA path (bod_pdq_*) against
goldens minted by executing the reference's offline_eval/pdq_data_holders.py + pdq.py verbatim
(tests/golden/make_pdq_golden.py).

Tolerances (floating point throughout; nothing here decides an index of the detection path):
  * regions of interest: exact integers;
  * heat maps: the oracle reproduces the reference bit for bit on the goldens; the CUDA maps must have the same
    zero pattern (the 0.0027 threshold) and agree within 2e-7 absolute (one float32 ulp below 1.0 — CUDA's
    erfc/exp/sin are within a few binary64 ulp, so a float32 rounding can flip);
  * loss sums: binary64 sums of binary32 terms here, float32 BLAS sums in the reference -> rtol 2e-4 against the
    goldens, rtol 1e-6 between CUDA and oracle;
  * qualities / cost tables: atol 1e-4 against the goldens; TP / FP / FN exact.
"""
import numpy as np
import pytest

from helpers import GOLDEN_DIR, pdq_golden_cases
from oracle import pdq as opdq

CASES = pdq_golden_cases()
HM_ATOL = 2e-7


def load(name):
    return np.load(f"{GOLDEN_DIR}/{name}.npz")


def corner_args(g):
    H, W = (int(v) for v in g["img_size"])
    out = []
    for b, c in zip(g["boxes"], g["covs"]):       # calc_heatmap :96-103
        out.append(((H, W), [b[1], b[0]], np.flipud(np.fliplr(c[0]))))
        out.append(((H, W), [H - (b[3] + 1), W - (b[2] + 1)], np.flipud(np.fliplr(c[1])).T))
    return out


def test_fixtures_present():
    assert len(CASES) >= 4


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_bvn_cdf_known_values():
    # closed forms: r = 0 -> product; h = k = 0 -> 1/4 + asin(r) / (2 pi); r = +-1 -> degenerate
    import math
    Phi = lambda x: 0.5 * math.erfc(-x / math.sqrt(2))  # noqa: E731
    for h, k in [(0.3, -1.2), (2.0, 2.5), (-3.0, 0.1)]:
        assert abs(opdq.bvn_cdf(h, k, 0.0) - Phi(h) * Phi(k)) < 1e-15
        assert abs(opdq.bvn_cdf(h, k, 1.0) - Phi(min(h, k))) < 1e-15
        assert abs(opdq.bvn_cdf(h, k, -1.0) - max(0.0, Phi(h) + Phi(k) - 1)) < 1e-15
    for r in [-0.99, -0.93, -0.8, -0.5, -0.2, 0.1, 0.29, 0.31, 0.74, 0.76, 0.92, 0.93, 0.999]:
        assert abs(opdq.bvn_cdf(0.0, 0.0, r) - (0.25 + math.asin(r) / (2 * math.pi))) < 2e-16 * 8, r
    # symmetry and the complement identity P(X<=h,Y<=k; r) = Phi(h) - P(X<=h, Y<=-k; -r)
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, k = rng.normal(size=2) * 2
        r = rng.uniform(-0.9999, 0.9999)
        a = opdq.bvn_cdf(h, k, r)
        assert abs(a - opdq.bvn_cdf(k, h, r)) < 1e-15
        assert abs(a - (Phi(h) - opdq.bvn_cdf(h, -k, -r))) < 1e-14, (h, k, r)


@pytest.mark.parametrize("name", CASES)
def test_oracle_roi_exact(name):
    g = load(name)
    rois = np.array([opdq.find_roi(*a) for a in corner_args(g)], np.int32).reshape(-1, 2, 4)
    assert np.array_equal(rois, g["rois"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_heatmaps_and_losses(name):
    g = load(name)
    H, W = (int(v) for v in g["img_size"])
    hm = opdq.heatmaps((H, W), g["boxes"], g["covs"])
    assert np.array_equal(hm > 0, g["heatmaps"] > 0)
    assert np.abs(hm - g["heatmaps"]).max() <= 6e-8          # bit-identical when minted; one ulp of slack across libm versions
    fg, bg, tot = opdq.losses(hm, g["gt_boxes"])
    np.testing.assert_allclose(fg, g["fg_loss"], rtol=2e-4, atol=1e-3)
    np.testing.assert_allclose(bg, g["bg_loss"], rtol=2e-4, atol=1e-3)
    spatial = np.exp((fg + bg) / g["num_fg"])
    np.testing.assert_allclose(spatial, g["spatial"], atol=1e-4)
    # whole-image background term against a direct evaluation
    direct = [(np.log((1 - m) + np.float32(1e-14)) * (m > 0)).astype(np.float64).sum() for m in hm]
    np.testing.assert_allclose(tot, direct, rtol=1e-6)


def test_oracle_rejects_what_the_reference_rejects():
    # the corner mean lies below the image: find_roi's row shift does not fit its own window (pdq_data_holders.py:163-172)
    with pytest.raises(ValueError):
        opdq.find_roi((64, 96), [80.0, 10.0], np.array([[4.0, 0.0], [0.0, 4.0]]))
    with pytest.raises(ValueError):
        opdq.single_heatmap((64, 96), [10.0, 10.0], np.array([[-1.0, 0.0], [0.0, 4.0]]))


def test_host_qualities_from_golden_losses():
    """pdq.py's host half (restated in bayes_od_rc_b200/pdq.py::_qual_img) fed with the reference's own loss sums."""
    from bayes_od_rc_b200 import pdq as ppdq
    for name in CASES:
        g = load(name)
        H, W = (int(v) for v in g["img_size"])
        gts = [ppdq.GroundTruthBox(b, l, (H, W)) for b, l in zip(g["gt_boxes"], g["gt_labels"])]
        dets = [ppdq.PBoxDet(c, b, cv) for c, b, cv in zip(g["cat_param"], g["boxes"], g["covs"])]
        assert [x.num_pixels for x in gts] == g["num_fg"].ravel().tolist()
        tot = [(np.log((1 - m) + np.float32(1e-14)) * (m > 0)).sum() for m in g["heatmaps"]]
        res, tables = ppdq._qual_img(gts, dets, [x.num_pixels for x in gts], g["fg_loss"], g["bg_loss"], tot)
        for k in ("overall", "spatial", "label"):
            np.testing.assert_allclose(tables[k], g["cost_" + k], atol=1e-6)
            np.testing.assert_allclose(res[k], float(g["res_" + k]), atol=1e-5)
        assert [res["TP"], res["FP"], res["FN"]] == g["res_counts"].tolist()


def test_host_rejects_non_box_ground_truth_and_handles_empty_images():
    """The evaluator's host half: masks that are not their box are refused (no silent approximation); images without
    objects or without detections never reach the GPU and follow pdq.py:351-368."""
    from bayes_od_rc_b200 import pdq as ppdq
    H, W = 64, 96
    gt = ppdq.GroundTruthBox([10, 12, 40, 50], 1, (H, W))
    gt.segmentation_mask = np.zeros((H, W), bool)
    gt.segmentation_mask[12:50, 10:40] = True
    assert ppdq._num_pixels(gt, (H, W)) == 30 * 38 == gt.num_pixels
    gt.segmentation_mask[0, 0] = True                      # a pixel outside the box
    with pytest.raises(ValueError):
        ppdq._num_pixels(gt, (H, W))
    gt.segmentation_mask[0, 0] = False
    gt.segmentation_mask[20, 20] = False                   # a hole
    with pytest.raises(ValueError):
        ppdq._num_pixels(gt, (H, W))
    big = ppdq.GroundTruthBox([0, 0, 50, 50], 0, (H, W))
    tiny = ppdq.GroundTruthBox([0, 0, 5, 50], 0, (H, W))     # too narrow to count (pdq.py:455-471)
    det = ppdq.PBoxDet(np.array([0.9, 0.1]), [1, 1, 20, 20], [np.eye(2), np.eye(2)])
    res, tables = ppdq._qual_img([big, tiny], [], [big.num_pixels, tiny.num_pixels], None, None, None)
    assert tables is None and res == {'overall': 0.0, 'spatial': 0.0, 'label': 0.0, 'TP': 0, 'FP': 0, 'FN': 1}
    res, _ = ppdq._qual_img([], [det, det], [], None, None, None)
    assert (res['TP'], res['FP'], res['FN']) == (0, 2, 0)
    clipped = ppdq.GroundTruthBox([80, 50, 120, 90], 0, (H, W))      # box reaching past the image: mask slicing clips it
    assert clipped.num_pixels == (96 - 80) * (64 - 50)


def test_det_instances_from_arrays_follow_compute_pdq():
    """compute_pdq.py:93-124 on the golden's saved arrays gives the boxes / corner covariances the golden was minted with."""
    from bayes_od_rc_b200 import pdq as ppdq
    for name in CASES:
        g = load(name)
        dets = ppdq.det_instances_from_arrays(g["means_vuhw"], g["covs_vuhw"], g["cat_param"], min_score=0.0)
        assert np.array_equal(np.array([d.box for d in dets]), g["boxes"])
        np.testing.assert_allclose(np.array([np.stack(d.covs) for d in dets]), g["covs"], rtol=1e-12)
        assert len(ppdq.det_instances_from_arrays(g["means_vuhw"], g["covs_vuhw"], g["cat_param"], min_score=2.0)) == 0
    assert ppdq.det_instances_from_arrays(np.zeros((0, 4, 1)), np.zeros((0, 4, 1)), np.zeros((0, 4, 1))) == []


# ------------------------------------------------------------------------------------------------ GPU: the product
@pytest.fixture(scope="module")
def engines():
    from bayes_od_rc_b200 import pdq as ppdq
    cache = {}

    def get(size):
        size = (int(size[0]), int(size[1]))
        if size not in cache:
            cache[size] = ppdq.PdqEngine(size)
        return cache[size]
    yield get
    for e in cache.values():
        e.close()


@pytest.mark.gpu
def test_gpu_bvn_cdf_matches_oracle(engines):
    rng = np.random.default_rng(1)
    n = 4000
    h, k = rng.normal(size=(2, n)) * 2.5
    r = np.concatenate([rng.uniform(-1, 1, n - 12), [0, 0.3, 0.75, 0.925, -0.925, 0.9999, -0.9999, 1, -1, 0.2999, -0.7499, 0.93]])
    got = engines((64, 96)).bvn_cdf(h, k, r)
    want = np.array([opdq.bvn_cdf(a, b, c) for a, b, c in zip(h, k, r)])
    assert np.abs(got - want).max() < 5e-15


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_heatmaps_match_reference(name, engines):
    g = load(name)
    hm = engines(g["img_size"]).heatmaps(g["boxes"], g["covs"])
    assert hm.shape == g["heatmaps"].shape
    assert np.array_equal(hm > 0, g["heatmaps"] > 0)
    assert np.abs(hm - g["heatmaps"]).max() <= HM_ATOL
    assert np.mean(hm == g["heatmaps"]) > 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_losses_and_pdq_match_reference(name, engines):
    from bayes_od_rc_b200 import pdq as ppdq
    g = load(name)
    H, W = (int(v) for v in g["img_size"])
    eng = engines((H, W))
    D, G = len(g["boxes"]), len(g["gt_boxes"])
    fgs, bgs, tot = eng.losses([0, D], g["boxes"], g["covs"], [0, G], g["gt_boxes"])
    np.testing.assert_allclose(fgs[0], g["fg_loss"], rtol=2e-4, atol=1e-3)
    np.testing.assert_allclose(bgs[0], g["bg_loss"], rtol=2e-4, atol=1e-3)
    ofg, obg, otot = opdq.losses(opdq.heatmaps((H, W), g["boxes"], g["covs"]), g["gt_boxes"])
    np.testing.assert_allclose(fgs[0], ofg, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(bgs[0], obg, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(tot, otot, rtol=1e-6, atol=1e-6)
    # the evaluator, fed like compute_pdq.py feeds the reference's (dense masks included)
    gts = []
    for b, l in zip(g["gt_boxes"], g["gt_labels"]):
        x = ppdq.GroundTruthBox(b, l, (H, W))
        x.segmentation_mask = np.zeros((H, W), bool)
        x.segmentation_mask[b[1]:b[3], b[0]:b[2]] = True
        gts.append(x)
    dets = [ppdq.PBoxDet(c, b, cv) for c, b, cv in zip(g["cat_param"], g["boxes"], g["covs"])]
    ev = ppdq.PDQ((H, W))
    ev._engine.close(); ev._engine = eng
    (res, tables), = ev.evaluate_images([(gts, dets)])
    for k in ("overall", "spatial", "label"):
        np.testing.assert_allclose(tables[k], g["cost_" + k], atol=1e-4)
        np.testing.assert_allclose(res[k], float(g["res_" + k]), atol=1e-4)
    assert [res["TP"], res["FP"], res["FN"]] == g["res_counts"].tolist()
    score = ev.score([(gts, dets)])
    assert abs(score - float(g["res_overall"]) / g["res_counts"].sum()) < 1e-4
    assert ev.get_assignment_counts() == tuple(g["res_counts"].tolist())


def _random_scene(rng, H, W, D, G, sig):
    boxes, covs = [], []
    for _ in range(D):
        h, w = rng.uniform(20, H * 0.6), rng.uniform(20, W * 0.6)
        y1, x1 = rng.uniform(0, H - h - 1), rng.uniform(0, W - w - 1)
        boxes.append([x1, y1, x1 + w, y1 + h])
        cs = []
        for _ in range(2):
            a = rng.normal(size=(2, 2))
            cs.append((a @ a.T + 0.3 * np.eye(2)) * rng.uniform(*sig))
        covs.append(cs)
    gt = []
    for _ in range(G):
        h, w = rng.uniform(15, H * 0.5), rng.uniform(15, W * 0.5)
        y1, x1 = rng.uniform(0, H - h), rng.uniform(0, W - w)
        gt.append([x1, y1, x1 + w, y1 + h])
    return np.array(boxes).astype(np.int32), np.array(covs), np.array(gt).astype(np.int32)


@pytest.mark.gpu
def test_gpu_full_size_batch_against_oracle(engines):
    """720x1280 (the size compute_pdq.py evaluates): dense maps and batched loss sums vs the oracle; a batch of
    images gives the same bits as one image at a time; images without detections or objects are legal."""
    rng = np.random.default_rng(7)
    H, W = 720, 1280
    eng = engines((H, W))
    scenes = [_random_scene(rng, H, W, D, G, sig) for D, G, sig in [(4, 3, (2, 30)), (3, 0, (50, 400)), (0, 2, (1, 2)), (5, 6, (0.05, 3))]]
    do, go = [0], [0]
    for b, c, g in scenes:
        do.append(do[-1] + len(b)); go.append(go[-1] + len(g))
    boxes = np.concatenate([s[0].reshape(-1, 4) for s in scenes]); covs = np.concatenate([s[1].reshape(-1, 2, 2, 2) for s in scenes])
    gts = np.concatenate([s[2].reshape(-1, 4) for s in scenes])
    fgs, bgs, tot = eng.losses(do, boxes, covs, go, gts)
    ms = eng.last_ms()
    assert ms["launches"] >= 3 and ms["table_floats"] > 0
    for i, (b, c, g) in enumerate(scenes):
        assert fgs[i].shape == (len(g), len(b))
        if len(b) == 0:
            continue
        hm = eng.heatmaps(b, c)
        ohm = opdq.heatmaps((H, W), b, c)
        assert np.array_equal(hm > 0, ohm > 0)
        assert np.abs(hm - ohm).max() <= HM_ATOL
        ofg, obg, otot = opdq.losses(ohm, g)
        np.testing.assert_allclose(fgs[i], ofg, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(bgs[i], obg, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(tot[do[i]:do[i + 1]], otot, rtol=1e-6, atol=1e-6)
        f1, b1, t1 = eng.losses([0, len(b)], b, c, [0, len(g)], g)
        assert np.array_equal(f1[0], fgs[i]) and np.array_equal(b1[0], bgs[i]) and np.array_equal(t1, tot[do[i]:do[i + 1]])


@pytest.mark.gpu
def test_gpu_evaluator_mixed_batch(engines):
    """PDQ.score over a batch mixing ordinary images with images that have no detections / no objects: same totals
    as scoring the images one at a time (add_img_eval), batch size irrelevant."""
    from bayes_od_rc_b200 import pdq as ppdq
    g = load("pdq_small")
    H, W = (int(v) for v in g["img_size"])
    gts = [ppdq.GroundTruthBox(b, l, (H, W)) for b, l in zip(g["gt_boxes"], g["gt_labels"])]
    dets = [ppdq.PBoxDet(c, b, cv) for c, b, cv in zip(g["cat_param"], g["boxes"], g["covs"])]
    matches = [(gts, dets), (gts, []), ([], dets), ([], []), (gts[:2], dets[1:4])]
    a = ppdq.PDQ((H, W), images_per_call=2)
    b = ppdq.PDQ((H, W), images_per_call=64)
    c = ppdq.PDQ((H, W))
    sa, sb = a.score(matches), b.score(matches)
    for m in matches:
        c.add_img_eval(*m)
    assert sa == sb == c.get_pdq_score()
    assert a.get_assignment_counts() == b.get_assignment_counts() == c.get_assignment_counts()
    tp, fp, fn = a.get_assignment_counts()
    assert fp >= len(dets) and fn >= sum(1 for x in gts if ppdq._is_gt_included(x, x.num_pixels))
    assert 0 < a.get_avg_spatial_score() <= 1 and 0 < a.get_avg_label_score() <= 1 and 0 < a.get_avg_overall_quality_score() <= 1


@pytest.mark.gpu
def test_gpu_heatmaps_into_device_tensor(engines):
    import torch
    g = load("pdq_small")
    eng = engines(g["img_size"])
    out = torch.full(g["heatmaps"].shape, -1.0, device="cuda")
    eng.heatmaps(g["boxes"], g["covs"], out=out)
    assert np.array_equal(out.cpu().numpy(), eng.heatmaps(g["boxes"], g["covs"]))


@pytest.mark.gpu
def test_gpu_roi_row_scan_matches_dense_scan(engines):
    """The ROI kernel settles the ends of each window row from the quadratic + the reference's predicate instead of scanning
    the whole 5-sigma window like find_roi (and the oracle) do: random corners with strong correlation, sub-pixel and
    image-sized sigmas and windows clipped by the image must give the same maps (a different ROI moves the replicated rows /
    columns of a corner map, so it cannot hide)."""
    rng = np.random.default_rng(123)
    H, W = 97, 143
    eng = engines((H, W))
    checked = 0
    for _ in range(25):
        boxes, covs = [], []
        for _ in range(8):
            x1, y1 = rng.integers(0, W - 12), rng.integers(0, H - 12)
            x2, y2 = rng.integers(x1 + 4, W), rng.integers(y1 + 4, H)
            cs = []
            for _ in range(2):
                s1, s2 = np.exp(rng.uniform(np.log(0.3), np.log(40), 2))
                r = rng.uniform(-0.97, 0.97)
                cs.append(np.array([[s1 * s1, r * s1 * s2], [r * s1 * s2, s2 * s2]]))
            boxes.append([x1, y1, x2, y2]); covs.append(cs)
        boxes, covs = np.array(boxes, np.int32), np.array(covs)
        try:
            ohm = opdq.heatmaps((H, W), boxes, covs)
        except ValueError:                      # a corner for which the reference raises: the product must refuse it too
            from bayes_od_rc_b200._cabi import BodError
            with pytest.raises(BodError):
                eng.heatmaps(boxes, covs)
            continue
        hm = eng.heatmaps(boxes, covs)
        assert np.array_equal(hm > 0, ohm > 0) and np.abs(hm - ohm).max() <= HM_ATOL
        checked += len(boxes)
    assert checked >= 100


@pytest.mark.gpu
def test_gpu_odd_width_uses_the_scalar_map_kernel(engines):
    rng = np.random.default_rng(11)
    H, W = 50, 70                                  # W % 4 != 0
    b, c, g = _random_scene(rng, H, W, 5, 3, (0.5, 20))
    hm = engines((H, W)).heatmaps(b, c)
    ohm = opdq.heatmaps((H, W), b, c)
    assert np.array_equal(hm > 0, ohm > 0) and np.abs(hm - ohm).max() <= HM_ATOL
    fgs, bgs, tot = engines((H, W)).losses([0, 5], b, c, [0, 3], g)
    ofg, obg, otot = opdq.losses(ohm, g)
    np.testing.assert_allclose(fgs[0], ofg, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(bgs[0], obg, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_gpu_errors(engines):
    from bayes_od_rc_b200._cabi import BodError
    eng = engines((64, 96))
    ok = np.array([[[4.0, 0], [0, 4.0]]] * 2)
    with pytest.raises(BodError, match="variance"):
        eng.heatmaps([[10, 10, 40, 40]], [np.array([[[-1.0, 0], [0, 4.0]], [[4.0, 0], [0, 4.0]]])])
    with pytest.raises(BodError, match="reference"):
        eng.heatmaps([[10, 90, 40, 40]], [ok])          # top-left corner below the image: find_roi raises in the reference
    assert eng.heatmaps(np.zeros((0, 4)), np.zeros((0, 8))).shape == (0, 64, 96)
    assert eng.heatmaps([[10, 10, 40, 40]], [ok]).max() > 0.5    # the context survives an error


def test_pdq_create_validates_before_touching_the_device():
    import ctypes as C
    from bayes_od_rc_b200 import _cabi
    lib = _cabi.load()
    h = C.c_void_p()
    for hh, ww in [(0, 10), (10, 0), (-1, 5), (40000, 40000)]:            # > 2^30 pixels: indices are 32-bit
        assert lib.bod_pdq_create(C.byref(h), 0, hh, ww) == -1
        assert b"bad argument" in lib.bod_pdq_last_error(None)
    assert lib.bod_pdq_create(None, 0, 8, 8) == -1
    # NULL contexts are refused, not dereferenced
    assert lib.bod_pdq_heatmaps(None, 0, None, None, None, 0) == -1
    assert lib.bod_pdq_losses(None, 0, None, None, None, None, None, None, None, None) == -1
    assert lib.bod_pdq_last_ms(None, None, None, None) == -1
    lib.bod_pdq_destroy(None)


def test_pdq_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    from bayes_od_rc_b200 import pdq as ppdq
    from bayes_od_rc_b200._cabi import BodError
    with pytest.raises(BodError):
        ppdq.PdqEngine((64, 96))
