"""Shared helpers of the parity tests."""
import glob
import json
import os

import numpy as np

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")

# north_star tolerance for fused means / covariances / class posteriors
RTOL, ATOL = 1e-4, 1e-5


def _all_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_cases():
    """Fixtures of bayes_od_inference / bayes_od_clustering (inference_utils.py)."""
    return [c for c in _all_cases() if not c.startswith(("val_", "pdq_", "writers_", "mue_"))]      # incl. the full-size "full_*" ones


def mue_golden_cases():
    """Fixtures of evaluation_utils_2d.py's entropy / MUE scoring (tests/golden/make_mue_golden.py)."""
    return [c for c in _all_cases() if c.startswith("mue_")]


def load_mue_golden(name):
    """-> dict with the arrays, meta, results and the gt / pred dict lists the reference functions take."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files if k not in ("meta", "results")}
    g["meta"] = json.loads(str(z["meta"])); g["results"] = json.loads(str(z["results"]))
    cats = g["meta"]["categories"]
    g["gt"] = [dict(name=f"frame_{int(i):04d}", category=cats[int(c)], bbox=[float(v) for v in b])
               for i, c, b in zip(g["gt_image"], g["gt_cat"], g["gt_box"])]
    g["pred"] = [dict(name=f"frame_{int(i):04d}", category=cats[int(c)], bbox=[float(v) for v in b])
                 for i, c, b in zip(g["pred_image"], g["pred_cat"], g["pred_box"])]
    return g


def pdq_golden_cases():
    """Fixtures of offline_eval/pdq_data_holders.py + pdq.py (tests/golden/make_pdq_golden.py)."""
    return [c for c in _all_cases() if c.startswith("pdq_")]


def val_golden_cases():
    """Fixtures of validation_utils.post_process_predictions."""
    return [c for c in _all_cases() if c.startswith("val_")]


def val_scaling_of(meta):
    """(scale_mode, shift, norm_hw, scale_hw) as validation_utils.py:54-66 derives them."""
    if meta["dataset_name"] == "kitti":
        return 1, (0, 0, 0, 0), tuple(float(x) for x in meta["image_shape"]), tuple(float(x) for x in meta["orig_size"])
    if meta["dataset_name"] == "coco":          # :60-66: shift by the padding, normalise by the unpadded size
        pad = [float(x) for x in meta["padding"]]
        shp = [int(meta["image_shape"][i]) - int(2 * pad[i]) for i in (0, 1)]
        return 2, tuple(pad), (float(shp[0]), float(shp[1])), tuple(float(x) for x in meta["orig_size"])
    return 0, (0, 0, 0, 0), (1.0, 1.0), (1.0, 1.0)


_FULL_CACHE = {}


def _regenerate_full_inputs(g):
    """Full-size fixtures (tests/golden/make_golden.py FULL_CASES) keep the generator arguments and a digest of the input
    bytes instead of the inputs: rebuild them exactly as the minting script did, or skip if this torch draws differently."""
    import hashlib

    import pytest
    import torch
    from bayes_od_rc_b200 import synthetic
    meta = g["meta"]
    spec = synthetic.SceneSpec(**meta["spec"])
    anchors = oracle.generate_anchors(spec.im_h, spec.im_w)          # bit-identical to the reference's (anchor_digests.json)
    if meta.get("kind") == "val":                                    # validation fixtures: one sample, no counts / covariances
        img = synthetic.make_image(spec, 0, torch.from_numpy(anchors), "cpu", with_counts=False)
        cls16, box16 = img["cls"][0].numpy().astype(np.float16), img["box"][0].numpy().astype(np.float16)
        h = hashlib.sha256()
        for arr in (anchors, cls16, box16):
            h.update(np.ascontiguousarray(arr).tobytes())
        if h.hexdigest() != meta["input_sha256"]:
            pytest.skip(f"{meta['case']}: the seeded generator produced different bytes here (torch {torch.__version__})")
        g.update(anchors=anchors, cls=cls16, box=box16)
        return
    img = synthetic.make_image(spec, 0, torch.from_numpy(anchors), "cpu", with_counts=True)
    cls16, box16, cov16 = (img[k].numpy().astype(np.float16) for k in ("cls", "box", "cov"))
    counts = img["counts"].numpy().astype(np.uint8)
    h = hashlib.sha256()
    for arr in (anchors, cls16, box16, cov16, counts):
        h.update(np.ascontiguousarray(arr).tobytes())
    if h.hexdigest() != meta["input_sha256"]:
        pytest.skip(f"{meta['case']}: the seeded generator produced different bytes here (torch {torch.__version__}); "
                    "the fixture only holds the outputs for its own inputs")
    g.update(anchors=anchors, cls=cls16, box=box16, cov=cov16, counts=counts)
    S, D = len(g["cnt_post"]), len(g["nms_indices"])
    g["iou_cols"] = np.unpackbits(g.pop("members"), axis=0, count=S).astype(np.float32).reshape(S, D)   # 1.0 = member


def load_golden(name):
    if name in _FULL_CACHE:
        return dict(_FULL_CACHE[name])
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(str(g["meta"]))
    if "input_sha256" in g["meta"]:
        _regenerate_full_inputs(g)
    g["cls"] = g["cls"].astype(np.float32)
    g["box"] = g["box"].astype(np.float32)
    for k in ("cov", "counts"):
        if k in g:
            g[k] = g[k].astype(np.float32)
    if "input_sha256" in g["meta"]:
        _FULL_CACHE[name] = dict(g)
    return g


def oracle_config_of(meta) -> oracle.OracleConfig:
    cfg = meta["cfg"]
    b, n = cfg["bayes_od_config"], cfg["nms_config"]
    sv = su = 1.0
    if meta["dataset_name"] == "kitti":   # inference_utils.py:151-152: int/int -> float64, cast to float32 at :159
        sv = float(np.float32(np.float64(meta["orig_size"][0]) / np.float64(meta["image_shape"][0])))
        su = float(np.float32(np.float64(meta["orig_size"][1]) / np.float64(meta["image_shape"][1])))
    return oracle.OracleConfig(
        use_full_covar=cfg["use_full_covar"], cov_layout=1 if meta["has_cov"] else 0,
        dirichlet_prior=b["dirichlet_prior"]["type"], gaussian_prior=b["gaussian_prior"]["type"],
        isotropic_variance=b["gaussian_prior"]["isotropic_variance"], ranking_method=b["ranking_method"],
        max_output_size=n["max_output_size"], iou_threshold=n["iou_threshold"], soft_nms_sigma=n["soft_nms_sigma"],
        scale_v=sv, scale_u=su)


def within_tol(got, ref, rtol=RTOL, atol=ATOL):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return np.abs(got - ref) <= atol + rtol * np.abs(ref)


def adjudicated_close(got32, ref32, ref64, rtol=RTOL, atol=ATOL):
    """SURVEY.md §7 hard part 3: an element passes if it is within tolerance of the
    binary32 reference, or no further from the binary64 value than the binary32
    reference itself is (plus tolerance).  Returns (ok_mask, plain_pass_fraction)."""
    got32 = np.asarray(got32, np.float64); ref32 = np.asarray(ref32, np.float64); ref64 = np.asarray(ref64, np.float64)
    plain = within_tol(got32, ref32, rtol, atol)
    adj = np.abs(got32 - ref64) <= np.abs(ref32 - ref64) + atol + rtol * np.abs(ref64)
    return plain | adj, float(plain.mean()) if plain.size else 1.0


def check_categorical_merge(g, nms_indices, member_bool, chosen, final_scores, final_counts, kl_slack=1e-5):
    """Validate the top-3-KL Dirichlet merge (inference_utils.py:333-354) against a
    golden fixture.  np.argpartition's order among EQUAL KL values is
    implementation-defined (introselect vs AVX-512 dispatch), and exact KL ties
    between different count vectors are common (e.g. [29,1,0..] vs [29,0,1..]), so:
      * clusters whose 3rd and 4th smallest KL differ: outputs must match the golden;
      * tied clusters: the picks must be a valid top-3 (their KLs are the three
        smallest up to float noise) and the outputs must equal mean/sum of the picks.
    Returns (n_unambiguous, n_tied)."""
    from scipy.stats import entropy
    cnt = g["cnt_post"]
    n_plain = n_tied = 0
    for d, c in enumerate(nms_indices):
        idx = np.flatnonzero(member_bool[d])
        if len(idx) <= 3:
            assert within_tol(final_counts[d], g["final_counts"][d]).all()
            assert within_tol(final_scores[d], g["final_scores"][d]).all()
            assert (chosen[d] == -1).all()
            n_plain += 1
            continue
        fs = cnt[idx] / cnt[idx].sum(1, keepdims=True)
        cs = np.repeat((cnt[c] / cnt[c].sum())[None], len(idx), 0)
        kl = entropy(cs.T, fs.T).astype(np.float64)               # as :339-344
        srt = np.sort(kl)
        picks = np.asarray(chosen[d])
        assert set(picks.tolist()) <= set(idx.tolist()) and len(set(picks.tolist())) == 3
        pk = np.sort(kl[np.searchsorted(idx, picks)])
        with np.errstate(invalid="ignore"):
            assert np.all((pk == srt[:3]) | (np.abs(pk - srt[:3]) <= kl_slack)), (d, pk, srt[:4])
        exp_counts = cnt[picks].sum(0)
        exp_scores = (cnt[picks] / cnt[picks].sum(1, keepdims=True)).mean(0)
        assert within_tol(final_counts[d], exp_counts).all()
        assert within_tol(final_scores[d], exp_scores).all()
        if np.isfinite(srt[3]) and srt[3] - srt[2] > kl_slack:                              # unambiguous -> must equal the golden
            assert within_tol(final_counts[d], g["final_counts"][d]).all(), d
            assert within_tol(final_scores[d], g["final_scores"][d]).all(), d
            n_plain += 1
        else:
            n_tied += 1
    return n_plain, n_tied
