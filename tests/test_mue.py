"""Entropy / minimum-uncertainty-error scoring (second half): the numpy restatement
(oracle/mue.py) against goldens minted from the reference's own functions (CPU), and the CUDA path behind the
drop-in (bayes_od_rc_b200/uncertainty.py) against both (GPU)."""
import copy

import numpy as np
import pytest

from helpers import load_mue_golden, mue_golden_cases
from oracle import mue as omue


def _with_scores(pred, ent):
    pr = copy.deepcopy(pred)
    for p, e in zip(pr, ent):
        p["entropy_score"] = e
    return pr


@pytest.mark.parametrize("name", mue_golden_cases())
def test_oracle_matches_reference_goldens(name):
    g = load_mue_golden(name)
    ge = omue.gaussian_entropy(g["covs"])
    assert np.allclose(ge, g["gaussian_entropy"], rtol=1e-12, atol=1e-12)
    ce = omue.categorical_entropy(g["params"])
    assert ce.dtype == np.float32 and np.array_equal(ce, g["categorical_entropy"])      # same numpy float32 operations
    for key, ref in g["results"].items():
        method, thr = key.split("@")
        ent = g["gaussian_entropy"] if method == "gaussian" else g["categorical_entropy"]
        mins, mean, cats, at = omue.evaluate_u_error(g["gt"], _with_scores(g["pred"], ent), [float(thr)])
        assert cats == ref["cats"]
        assert np.allclose(mins, ref["min_u_errors"], rtol=0, atol=1e-15) and abs(mean - ref["mean"]) < 1e-15
        assert np.allclose(at, ref["scores_at_min"], rtol=0, atol=0)


def test_there_are_mue_goldens():
    assert len(mue_golden_cases()) >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("name", mue_golden_cases())
def test_gpu_entropies_and_mue_match_reference(name):
    from bayes_od_rc_b200 import uncertainty as fast
    g = load_mue_golden(name)
    ge = fast.gaussian_entropies(g["covs"])
    # binary32 determinant (numpy runs LAPACK's single-precision LU on the binary32 covariances) and logf: a few ulp of
    # binary32 in 0.5 log(det), i.e. ~1e-6 in an entropy of ~15
    assert np.allclose(ge, g["gaussian_entropy"], rtol=0, atol=4e-6)
    ce = fast.categorical_entropies(g["params"])
    assert ce.dtype == np.float32 and np.allclose(ce, g["categorical_entropy"], rtol=2e-6, atol=2e-6)   # logf vs numpy's log
    assert abs(fast.compute_gaussian_entropy_np(g["covs"][0]) - g["gaussian_entropy"][0]) < 4e-6
    assert abs(fast.compute_categorical_entropy_np(g["params"][0]) - g["categorical_entropy"][0]) < 2e-6
    # the curve on the REFERENCE's entropies (ranking decided by exactly the same keys): exact counts => exact values
    for key, ref in g["results"].items():
        method, thr = key.split("@")
        ent = g["gaussian_entropy"] if method == "gaussian" else g["categorical_entropy"]
        pr = _with_scores(g["pred"], ent)
        mins, mean, cats, at = fast.evaluate_u_error(g["gt"], pr, [float(thr)])
        assert cats == ref["cats"]
        assert np.allclose(mins, ref["min_u_errors"], rtol=0, atol=1e-15) and abs(mean - ref["mean"]) < 1e-15
        assert np.array_equal(np.asarray(at), np.asarray(ref["scores_at_min"]))
        # and one category through compute_mu_error against the restatement, two thresholds at once
        cat = ref["cats"][0]
        gsub = [x for x in g["gt"] if x["category"] == cat]
        psub = [x for x in pr if x["category"] == cat]
        if psub:
            names = {}
            for x in gsub + psub:
                names.setdefault(x["name"], len(names))
            mo, flat, ranking, _ = omue.mu_error(np.asarray([x["bbox"] for x in psub]), [x["entropy_score"] for x in psub],
                                                 [names[x["name"]] for x in psub], np.asarray([x["bbox"] for x in gsub]).reshape(-1, 4),
                                                 [names[x["name"]] for x in gsub], [0.5, 0.75])
            ranked = np.asarray([x["entropy_score"] for x in psub])[ranking]
            if flat < len(ranked):
                m, s = fast.compute_mu_error(gsub, psub, [0.5, 0.75])
                assert abs(m - mo) < 1e-15 and s == ranked[flat]
            else:       # the reference indexes its score list with the flat arg-min of the [n, T] matrix (:211-213): IndexError there too
                with pytest.raises(IndexError):
                    fast.compute_mu_error(gsub, psub, [0.5, 0.75])


@pytest.mark.gpu
def test_gpu_mue_large_random():
    """100 k predictions over 2 000 images, one category: the curve kernel's chunked scan, images without ground
    truth, ties in the ranking."""
    from bayes_od_rc_b200 import uncertainty as fast
    rng = np.random.default_rng(7)
    n_img, n = 2000, 100000
    gt, pred = [], []
    for im in range(n_img):
        if im % 7 == 0:
            continue
        for _ in range(int(rng.integers(1, 9))):
            x1, y1 = rng.uniform(0, 1000), rng.uniform(0, 600)
            gt.append(dict(name=str(im), category="car", bbox=[x1, y1, x1 + rng.uniform(20, 200), y1 + rng.uniform(20, 200)]))
    gtb = {}
    for x in gt:
        gtb.setdefault(x["name"], []).append(x["bbox"])
    ents = np.round(rng.uniform(0, 3, n), 2)                       # many ties
    for i in range(n):
        im = str(int(rng.integers(0, n_img)))
        if im in gtb and rng.uniform() < 0.6:
            b = np.asarray(gtb[im][int(rng.integers(0, len(gtb[im])))]) + rng.normal(0, 8, 4)
        else:
            x1, y1 = rng.uniform(0, 1000), rng.uniform(0, 600)
            b = np.asarray([x1, y1, x1 + rng.uniform(20, 200), y1 + rng.uniform(20, 200)])
        pred.append(dict(name=im, category="car", bbox=[float(v) for v in b], entropy_score=float(ents[i])))
    m, s = fast.compute_mu_error(gt, pred, [0.5])
    names = {}
    for x in gt + pred:
        names.setdefault(x["name"], len(names))
    mo, flat, ranking, _ = omue.mu_error(np.asarray([x["bbox"] for x in pred]), ents, [names[x["name"]] for x in pred],
                                         np.asarray([x["bbox"] for x in gt]), [names[x["name"]] for x in gt], [0.5])
    assert abs(m - mo) < 1e-15 and s == ents[ranking][flat]
