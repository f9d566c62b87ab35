#!/usr/bin/env python
"""Mint the golden fixtures of tests/golden/*.npz.

Runs in the AUTHORING container only (needs /root/reference): imports the
reference's own inference_utils.py / box_utils.py / fpn_anchor_generator.py
over tests/golden/tf_numpy_shim.py and executes

    bayes_od_inference(model, sample_dict, bayes_od_config, nms_config, ...)   # inference_utils.py:13-217
    bayes_od_clustering(counts, means, covs, nms_indices, iou_mat, thr)        # inference_utils.py:285-364
    FpnAnchorGenerator.generate_anchors(...)                                   # fpn_anchor_generator.py:21-59

on small synthetic head outputs, with the (unseeded) categorical draws injected.
Inputs are stored as float16 (they are exactly representable, every consumer
upcasts to float32), outputs as float32.

    python tests/golden/make_golden.py            # rewrites every fixture
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_numpy_shim as shim  # noqa: E402
from bayes_od_rc_b200 import synthetic  # noqa: E402

BDD_TEST_CFG = dict(  # retinanet_bdd_covar.yaml:117-144
    use_full_covar=True,
    nms_config=dict(max_output_size=100, iou_threshold=0.5, soft_nms_sigma=0.5),
    bayes_od_config=dict(ranking_method='score', dirichlet_prior=dict(type='non_informative'),
                         gaussian_prior=dict(type='isotropic', isotropic_variance=100000.0), fusion_method='none'))

CASES = {
    # name: (SceneSpec kwargs, overrides)
    "bdd_covar_k8":   (dict(im_h=96, im_w=160, N=10, K=8, g_min=4, g_max=6, box_hi=90., config_id=11), {}),
    "bdd_kendall_k8": (dict(im_h=64, im_w=96, N=6, K=8, g_min=3, g_max=4, box_hi=60., config_id=12), dict(use_full_covar=False)),
    "bdd_covar_k11":  (dict(im_h=64, im_w=96, N=5, K=11, g_min=3, g_max=4, box_hi=60., config_id=13), {}),
    "kitti_k4_n8":    (dict(im_h=64, im_w=128, N=8, K=4, g_min=3, g_max=5, box_hi=60., config_id=14),
                       dict(dataset_name='kitti', orig_size=(47, 94))),
    "no_cov_head":    (dict(im_h=64, im_w=96, N=6, K=8, g_min=3, g_max=4, box_hi=60., config_id=15), dict(drop_cov=True)),
    # gaussian_prior 'None' cannot be minted: the reference itself raises at
    # inference_utils.py:205 (tf.squeeze(axis=2) of the un-expanded [S,4] means).
    "no_dirichlet":   (dict(im_h=64, im_w=96, N=4, K=8, g_min=3, g_max=4, box_hi=60., config_id=16),
                       dict(dirichlet='None')),
    "joint_entropy":  (dict(im_h=64, im_w=96, N=4, K=8, g_min=3, g_max=4, box_hi=60., config_id=17),
                       dict(ranking='joint_entropy')),
    "max_out_7":      (dict(im_h=64, im_w=96, N=4, K=8, g_min=3, g_max=4, box_hi=60., config_id=18), dict(max_output_size=7)),
    "single_survivor": (dict(im_h=64, im_w=96, N=4, K=8, g_min=3, g_max=4, box_hi=60., config_id=19), dict(force_survivors=1)),
    "no_survivor":    (dict(im_h=64, im_w=96, N=4, K=8, g_min=3, g_max=4, box_hi=60., config_id=20), dict(force_survivors=0)),
}


def f16_exact(t: torch.Tensor) -> np.ndarray:
    return t.numpy().astype(np.float16)


def ids_from_counts(counts: np.ndarray, T: int) -> np.ndarray:
    """[A,K] integer counts (rows sum to T) -> [T,A] class ids whose one-hot sum gives them back."""
    A, K = counts.shape
    ids = np.zeros((T, A), np.int32)
    for a in range(A):
        ids[:, a] = np.repeat(np.arange(K), counts[a].astype(np.int64))
    return ids


def run_case(name, spec_kw, ov, iu, bu, ag, cs, Categorical, store_inputs=True):
    spec = synthetic.SceneSpec(**spec_kw)
    # anchors from the REFERENCE's generator (fpn_anchor_generator.py:21-59), P3->P7
    gen = ag.FpnAnchorGenerator(dict(aspect_ratios=[[1.0, 1.0], [1.0, 2.0], [2.0, 1.0]], scales=[1.0, 1.26, 1.59]))
    image_norm = np.zeros((spec.im_h, spec.im_w, 3), np.float32)
    anchors = np.concatenate([np.asarray(gen.generate_anchors(shim._t(np.asarray(image_norm.shape, np.int32)), l))
                              for l in [3, 4, 5, 6, 7]], axis=0).astype(np.float32)
    img = synthetic.make_image(spec, 0, torch.from_numpy(anchors), "cpu", with_counts=True)
    cls16, box16, cov16 = f16_exact(img["cls"]), f16_exact(img["box"]), f16_exact(img["cov"])
    cls, box, cov = cls16.astype(np.float32), box16.astype(np.float32), cov16.astype(np.float32)
    counts = img["counts"].numpy().astype(np.float32)
    K = spec.K
    if "force_survivors" in ov:                     # edge cases: S = 0 / 1
        bg = np.zeros_like(counts); bg[:, K - 1] = 30.0
        keep = np.flatnonzero(np.argmax(counts, 1) != K - 1)[:ov["force_survivors"]]
        bg[keep] = counts[keep]
        counts = bg

    cfg = json.loads(json.dumps(BDD_TEST_CFG))
    if "use_full_covar" in ov: cfg["use_full_covar"] = ov["use_full_covar"]
    if "dirichlet" in ov: cfg["bayes_od_config"]["dirichlet_prior"]["type"] = ov["dirichlet"]
    if "gaussian" in ov: cfg["bayes_od_config"]["gaussian_prior"]["type"] = ov["gaussian"]
    if "ranking" in ov: cfg["bayes_od_config"]["ranking_method"] = ov["ranking"]
    if "max_output_size" in ov: cfg["nms_config"]["max_output_size"] = ov["max_output_size"]
    dataset_name = ov.get("dataset_name", "bdd")
    orig = ov.get("orig_size", (spec.im_h, spec.im_w))

    pred = {cs.ANCHORS_CLASS_PREDICTIONS_KEY: shim._t(cls), cs.ANCHORS_BOX_PREDICTIONS_KEY: shim._t(box)}
    if not ov.get("drop_cov"):
        pred[cs.ANCHORS_COVAR_PREDICTIONS_KEY] = shim._t(cov)
    model = lambda image, train_val_test='testing': pred   # noqa: E731  (stands in for RetinaNetModel.call)
    sample_dict = {cs.IMAGE_NORMALIZED_KEY: shim._t(image_norm[None]),
                   cs.ANCHORS_KEY: shim._t(anchors[None]),
                   cs.ORIGINAL_IM_SIZE_KEY: shim._t(np.asarray([[orig[0], orig[1], 3]], np.int32))}
    Categorical.forced_samples = ids_from_counts(counts, 30)
    out = iu.bayes_od_inference(model, sample_dict, cfg["bayes_od_config"], cfg["nms_config"],
                                use_full_covar=cfg["use_full_covar"], dataset_name=dataset_name)
    cnt_post, mu_post, sig_post, nms_idx, iou_mat = [np.asarray(o) for o in out]
    res = dict(cnt_post=cnt_post, mu_post=mu_post, sig_post=sig_post, nms_indices=nms_idx.astype(np.int32),
               iou_cols=iou_mat[:, nms_idx] if iou_mat.size else np.zeros((0, 0), np.float32))
    if mu_post.size > 0:                                            # run_inference.py:147-149
        fs, fm, fc, fn = iu.bayes_od_clustering(cnt_post, mu_post, sig_post, nms_idx, iou_mat,
                                                affinity_threshold=cfg["nms_config"]["iou_threshold"])
        res.update(final_scores=fs, final_means=fm, final_covs=fc, final_counts=fn)
        for k in ("final_scores", "final_means", "final_covs", "final_counts"):
            assert res[k].dtype == np.float32, (k, res[k].dtype)
    meta = dict(case=name, spec=spec_kw, cfg=cfg, dataset_name=dataset_name, orig_size=list(orig),
                image_shape=[spec.im_h, spec.im_w], has_cov=not ov.get("drop_cov", False),
                numpy=np.__version__, generator="tests/golden/make_golden.py over tf_numpy_shim; reference sources executed verbatim")
    if store_inputs:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), anchors=anchors,
                            cls=cls16, box=box16, cov=cov16, counts=counts.astype(np.uint8), **res)
    else:
        import hashlib
        hsh = hashlib.sha256()
        for arr in (anchors, cls16, box16, cov16, counts.astype(np.uint8)):
            hsh.update(np.ascontiguousarray(arr).tobytes())
        meta["input_sha256"] = hsh.hexdigest()
        meta["inputs"] = "regenerate: synthetic.make_image(SceneSpec(**spec), 0, anchors, 'cpu', with_counts=True), float16-rounded"
        res.pop("iou_cols")                                     # [S, D] float32: replaced by the membership bits below
        res["members"] = np.packbits(iou_mat[:, nms_idx] > cfg["nms_config"]["iou_threshold"], axis=0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), **res)
    S, D = len(cnt_post), len(nms_idx)
    members = (iou_mat[:, nms_idx] > 0.5).sum(0) if S else []
    print(f"{name:16s} A={len(anchors)} S={S} D={D} members(min/max)={min(members, default=0)}/{max(members, default=0)}")


# Full-size images (BASELINE.json shapes).  Inputs are NOT stored (hundreds of MB): the fixture keeps the generator
# arguments and a SHA-256 of the exact input bytes; tests regenerate them with the same seeded generator
# (bayes_od_rc_b200/synthetic.py on the CPU) and skip if the digest differs (another torch RNG).
FULL_CASES = {
    "full_bdd_covar_k8": (dict(im_h=720, im_w=1280, N=10, K=8, config_id=31), {}),
    "full_kitti_k4_n20": (dict(im_h=375, im_w=1242, N=20, K=4, config_id=32), dict(dataset_name='kitti', orig_size=(375, 1242))),
    "full_bdd_kendall_k11": (dict(im_h=720, im_w=1280, N=10, K=11, config_id=33), dict(use_full_covar=False)),
    "full_bdd_entropy_k8": (dict(im_h=720, im_w=1280, N=6, K=8, config_id=34), dict(ranking='joint_entropy')),
}


VAL_CASES = {
    # validation_utils.post_process_predictions (validation_utils.py:10-77): name -> (SceneSpec kwargs, overrides)
    "val_bdd_k8":   (dict(im_h=96, im_w=160, N=2, K=8, g_min=4, g_max=6, box_hi=90., config_id=21), {}),
    "val_kitti_k4": (dict(im_h=64, im_w=128, N=2, K=4, g_min=3, g_max=5, box_hi=60., config_id=22),
                     dict(dataset_name='kitti', orig_size=(47, 94))),
    "val_none_k8":  (dict(im_h=64, im_w=96, N=2, K=8, g_min=3, g_max=4, box_hi=60., config_id=23), dict(all_background=True)),
    # the coco branch (validation_utils.py:60-66): corners shifted by the applied padding, normalised by the unpadded size
    "val_coco_k8":  (dict(im_h=96, im_w=160, N=2, K=8, g_min=4, g_max=6, box_hi=90., config_id=26),
                     dict(dataset_name='coco', orig_size=(61, 113), padding=(8.0, 12.0, 8.0, 12.0))),
    # full size: outputs + input digest only (see FULL_CASES)
    "val_full_bdd_k8": (dict(im_h=720, im_w=1280, N=2, K=8, config_id=24), dict(digest_only=True)),
    "val_full_kitti_k4": (dict(im_h=375, im_w=1242, N=2, K=4, config_id=25),
                          dict(dataset_name='kitti', orig_size=(370, 1224), digest_only=True)),
}


def run_val_case(name, spec_kw, ov, vu, ag, cs):
    spec = synthetic.SceneSpec(**spec_kw)
    gen = ag.FpnAnchorGenerator(dict(aspect_ratios=[[1.0, 1.0], [1.0, 2.0], [2.0, 1.0]], scales=[1.0, 1.26, 1.59]))
    image_norm = np.zeros((spec.im_h, spec.im_w, 3), np.float32)
    anchors = np.concatenate([np.asarray(gen.generate_anchors(shim._t(np.asarray(image_norm.shape, np.int32)), l))
                              for l in [3, 4, 5, 6, 7]], axis=0).astype(np.float32)
    img = synthetic.make_image(spec, 0, torch.from_numpy(anchors), "cpu", with_counts=False)
    cls16, box16 = f16_exact(img["cls"][0]), f16_exact(img["box"][0])          # one sample: validation runs without dropout
    if ov.get("all_background"):
        cls16 = cls16.copy(); cls16[:, -1] = np.float16(20.0)
    cls, box = cls16.astype(np.float32), box16.astype(np.float32)
    dataset_name = ov.get("dataset_name", "bdd")
    orig = ov.get("orig_size", (spec.im_h, spec.im_w))
    pred = {cs.ANCHORS_CLASS_PREDICTIONS_KEY: shim._t(cls[None]), cs.ANCHORS_BOX_PREDICTIONS_KEY: shim._t(box[None])}
    sample_dict = {cs.IMAGE_NORMALIZED_KEY: shim._t(image_norm[None]), cs.ANCHORS_KEY: shim._t(anchors[None]),
                   cs.ORIGINAL_IM_SIZE_KEY: shim._t(np.asarray([[orig[0], orig[1], 3]], np.int32))}
    if "padding" in ov:
        sample_dict[cs.IMAGE_PADDING_KEY] = shim._t(np.asarray([ov["padding"]], np.float32))
    classes_out, corners_out = vu.post_process_predictions(sample_dict, pred, dataset_name=dataset_name)
    classes_out, corners_out = np.asarray(classes_out, np.float32), np.asarray(corners_out, np.float32)
    meta = dict(case=name, spec=spec_kw, dataset_name=dataset_name, orig_size=list(orig), image_shape=[spec.im_h, spec.im_w],
                padding=list(ov.get("padding", ())),
                numpy=np.__version__, generator="tests/golden/make_golden.py over tf_numpy_shim; "
                "validation_utils.post_process_predictions executed verbatim")
    if ov.get("digest_only"):
        import hashlib
        hsh = hashlib.sha256()
        for arr in (anchors, cls16, box16):
            hsh.update(np.ascontiguousarray(arr).tobytes())
        meta["input_sha256"] = hsh.hexdigest()
        meta["kind"] = "val"
        meta["inputs"] = "regenerate: synthetic.make_image(SceneSpec(**spec), 0, anchors, 'cpu', with_counts=False), sample 0, float16-rounded"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta),
                            classes_out=classes_out.reshape(-1, spec.K), corners_out=corners_out.reshape(-1, 4))
    else:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), anchors=anchors, cls=cls16, box=box16,
                            classes_out=classes_out.reshape(-1, spec.K), corners_out=corners_out.reshape(-1, 4))
    print(f"{name:16s} A={len(anchors)} D={len(classes_out)}")


def main():
    iu, bu, ag, cs, Categorical = shim.load_reference()
    only_val = "--val-only" in sys.argv or "--full" in sys.argv
    if not only_val:
        for name, (spec_kw, ov) in CASES.items():
            run_case(name, spec_kw, ov, iu, bu, ag, cs, Categorical)
    if not only_val or "--full" in sys.argv:
        for name, (spec_kw, ov) in FULL_CASES.items():
            run_case(name, spec_kw, ov, iu, bu, ag, cs, Categorical, store_inputs=False)
    import importlib
    vu = importlib.import_module("src.retina_net.experiments.validation_utils")
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for name, (spec_kw, ov) in VAL_CASES.items():
        if only and name not in only:
            continue
        run_val_case(name, spec_kw, ov, vu, ag, cs)


if __name__ == "__main__":
    main()
