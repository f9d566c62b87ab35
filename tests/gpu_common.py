"""Helpers of the GPU parity tests: run the CUDA path through the C ABI and
compare every stage with the CPU oracle on the same inputs."""
import numpy as np
import torch

import oracle
from bayes_od_rc_b200 import _cabi
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine


def engine_config_from_oracle(oc: oracle.OracleConfig, **kw) -> BayesODConfig:
    return BayesODConfig(
        use_full_covar=oc.use_full_covar, cov_layout=oc.cov_layout, dirichlet_prior=oc.dirichlet_prior,
        gaussian_prior=oc.gaussian_prior, isotropic_variance=oc.isotropic_variance, ranking_method=oc.ranking_method,
        max_output_size=oc.max_output_size, iou_threshold=oc.iou_threshold, soft_nms_sigma=oc.soft_nms_sigma,
        scale_v=oc.scale_v, scale_u=oc.scale_u, cov_calibration=oc.cov_calibration, num_draws=oc.num_draws,
        seed=oc.seed, image_id_base=oc.image_id_base, score_threshold=oc.score_threshold,
        pre_nms_top_k=oc.pre_nms_top_k, **kw)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(got, ref, what):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.dtype == np.float32:
        same = (bits(got) == bits(ref)) | (np.isnan(got) & np.isnan(ref))
    else:
        same = got == ref
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {same.size} elements differ; first at {i}: got {got[i]!r} ref {ref[i]!r}")


def run_gpu_batch(oc, cls, box, cov, anchors, counts, emit_probs=True, via_host=False, **engine_kw):
    """cls [B,N,A,K] ... numpy.  Returns (engine, results)."""
    B, N, A, K = cls.shape
    cfg = engine_config_from_oracle(oc, emit_probs=emit_probs, **engine_kw)
    eng = BayesODEngine(B, N, A, K, cfg, device=0)
    if via_host:
        res = eng.run_host(cls, box, cov if oc.cov_layout else None, anchors, counts)
    else:
        dev = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()   # noqa: E731
        t = [dev(cls), dev(box), dev(cov) if oc.cov_layout else None, dev(anchors), dev(counts)]
        torch.cuda.synchronize()
        eng.run(*t, stream=torch.cuda.current_stream().cuda_stream)
        res = eng.fetch()
    return eng, res


def compare_image_with_oracle(eng, res, b, r: "oracle.ImageResult", K, check_probs=True, exact_fused=True):
    """Bit-exact comparison of every stage of image b with the oracle result r."""
    S, D = len(r.keep), len(r.nms_indices)
    assert int(res.num_survivors[b]) == S, (int(res.num_survivors[b]), S)
    assert int(res.num_dets[b]) == D, (int(res.num_dets[b]), D)
    sv = eng.survivors(b)
    assert_bit_equal(sv["anchor_idx"], r.keep, "kept anchor indices")
    assert_bit_equal(sv["counts"], r.cnt_post, "dirichlet posterior counts")
    assert_bit_equal(sv["means"], r.mu_post, "posterior means")
    assert_bit_equal(sv["covs"], r.sig_post, "posterior covariances")
    assert_bit_equal(sv["scores"], r.score, "ranking scores")
    assert_bit_equal(sv["corners"], r.corners, "corners")
    assert_bit_equal(res.nms_indices[b, :D], r.nms_indices, "nms_indices")
    assert (res.nms_indices[b, D:] == -1).all()
    assert_bit_equal(res.centre_scores[b, :D], r.nms_scores, "soft-NMS scores at selection")
    assert_bit_equal(res.centre_anchor_idx[b, :D], r.keep[r.nms_indices], "centre anchor indices")
    if D:
        mask = eng.members(b, S, D)
        assert_bit_equal(mask, r.mask, "cluster membership bitmasks")
    if exact_fused:
        assert_bit_equal(res.means[b, :D], r.final_means, "fused means")
        assert_bit_equal(res.covs[b, :D], r.final_covs, "fused covariances")
        assert_bit_equal(res.cat_param[b, :D], r.final_scores, "fused class scores")
        assert_bit_equal(res.cat_count[b, :D], r.final_counts, "fused class counts")
    # padding rows are zero
    assert not res.means[b, D:].any() and not res.covs[b, D:].any()
    assert not res.cat_param[b, D:].any() and not res.cat_count[b, D:].any()
    if check_probs and r.probs is not None:
        p = eng.probs(b)
        err = np.abs(p - r.probs).max()
        assert err <= 2e-6, f"mean class probabilities differ by {err}"      # fast exp vs correctly rounded exp
