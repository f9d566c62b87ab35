"""Drop-in for the reference's ``src/retina_net/experiments/inference_utils.py``.

Same two names, same argument meaning, same return structure, so that
``run_inference.py:137-161`` and everything downstream (the .npy / json writers,
the offline MUE / AP / PDQ scripts) run unchanged:

    bayes_od_inference(model, sample_dict, bayes_od_config, nms_config,
                       use_full_covar=False, dataset_name='bdd')           # reference :13-217
        -> (dirichlit_posterior_count [S,K], gaussian_posterior_means [S,4,1],
            gaussian_posterior_covs [S,4,4], nms_indices [D], predicted_boxes_iou_mat)
    bayes_od_clustering(predicted_boxes_class_counts, predicted_boxes_means, predicted_boxes_covs,
                        cluster_centers, affinity_matrix, affinity_threshold=0.7)   # reference :285-364
        -> (final_box_class_scores [D,K], final_box_means [D,4,1],
            final_box_covs [D,4,4], final_box_class_counts [D,K])    float32

How it differs inside: the model call stays with the caller's framework (TF2 in
the reference); its three prediction tensors are handed to the CUDA library
through DLPack without a copy, and ALL of the post-head work — including the
clustering that the reference does on the host — runs on the GPU in one
``bod_run``.  The fifth return value, which in the reference is the dense
[S,S] IoU matrix (100 MB at S=5000) pulled to the host only so that
``bayes_od_clustering`` can read D of its columns, is replaced by a small
``FusedClusters`` token that carries the already fused detections;
``bayes_od_clustering`` recognises it.  Given a real ndarray affinity matrix
instead (any other caller), ``bayes_od_clustering`` runs the fusion kernel on the
caller's arrays (``bod_cluster_host``).  There is no CPU implementation here.

Install it over the reference with::

    import bayes_od_rc_b200.inference_utils as fast
    from src.retina_net.experiments import inference_utils
    inference_utils.bayes_od_inference = fast.bayes_od_inference
    inference_utils.bayes_od_clustering = fast.bayes_od_clustering

(see INTEGRATION.md).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .engine import BayesODConfig, BayesODEngine

# prediction / sample dict keys, src/core/constants.py:48-63
IMAGE_NORMALIZED_KEY = 'image_normalized'
ORIGINAL_IM_SIZE_KEY = 'im_size'
ANCHORS_KEY = 'anchors'
ANCHORS_BOX_PREDICTIONS_KEY = 'anchors_box_predictions'
ANCHORS_COVAR_PREDICTIONS_KEY = 'anchors_box_covar_predictions'
ANCHORS_CLASS_PREDICTIONS_KEY = 'anchors_class_predictions'


class HostTensor(np.ndarray):
    """ndarray with the ``.numpy()`` method run_inference.py:141-145 calls on every output."""

    def numpy(self):
        return np.asarray(self)


def _ht(a) -> HostTensor:
    return np.ascontiguousarray(a).view(HostTensor)


class FusedClusters:
    """Stands in for ``predicted_boxes_iou_mat``: the detections the GPU already
    fused with affinity threshold ``threshold`` for the centres ``nms_indices``."""

    def __init__(self, scores, means, covs, counts, nms_indices, threshold, num_survivors):
        self.scores, self.means, self.covs, self.counts = scores, means, covs, counts
        self.nms_indices, self.threshold, self.num_survivors = nms_indices, threshold, num_survivors
        self.shape = (num_survivors, num_survivors)       # what the dense matrix would have been
        self.size = num_survivors * num_survivors

    def numpy(self):
        return self


_engines = {}


_PER_RUN = ("seed", "image_id_base", "scale_v", "scale_u")     # set per call on a cached engine, not part of its identity


def _engine(B, N, A, K, cfg: BayesODConfig, device=0) -> BayesODEngine:
    """One engine (workspace, streams, events) per shape and configuration; the sampler stream and the KITTI
    scale follow the image, so they are applied per call instead of keying the cache (a caller that numbers its
    images would otherwise build a context per image)."""
    key = (B, N, A, K, device, tuple(sorted((k, v) for k, v in cfg.__dict__.items() if k not in _PER_RUN)))
    eng = _engines.get(key)
    if eng is None:
        if len(_engines) > 8:                      # shapes rarely change; do not hoard workspaces
            _engines.popitem()[1].close()
        eng = _engines[key] = BayesODEngine(B, N, A, K, cfg, device)
    eng.set_sampler_stream(cfg.seed, cfg.image_id_base)
    eng.set_image_scale(cfg.scale_v, cfg.scale_u)
    return eng


def _shape(x):
    return tuple(int(d) for d in x.shape)


_DL_CPU = (1, 3)            # kDLCPU, kDLCUDAHost


def _is_host(x) -> bool:
    """Host-resident producer output: numpy, or any tensor that says so through DLPack (tf.data / CPU tensors)."""
    if isinstance(x, np.ndarray):
        return True
    if hasattr(x, "is_cuda"):
        return not x.is_cuda
    if hasattr(x, "__dlpack_device__"):
        return int(x.__dlpack_device__()[0]) in _DL_CPU
    return False


def _to_numpy(x, shape=None):
    """float32 ndarray of a host tensor (numpy, torch CPU, anything with __dlpack__ / __array__)."""
    if x is None:
        return None
    if not isinstance(x, np.ndarray):
        if hasattr(x, "detach"):
            x = x.detach().cpu().numpy()
        elif hasattr(x, "__dlpack__") and hasattr(x, "__dlpack_device__"):
            x = np.from_dlpack(x)
        elif hasattr(x, "numpy"):
            x = x.numpy()
    a = np.ascontiguousarray(np.asarray(x), np.float32)
    return a.reshape(shape) if shape is not None else a


_image_counter = [0]        # images seen by bayes_od_inference: the default Philox stream id of each call


def bayes_od_inference(model, sample_dict, bayes_od_config, nms_config, use_full_covar=False, dataset_name='bdd',
                       counts=None, seed=1234, image_id=None, device=0):
    """See module docstring.  Extra keyword arguments (not in the reference):
    ``counts`` [A,K] injects the categorical draw counts (the reference draws them
    unseeded, inference_utils.py:37-46); otherwise the in-kernel Philox sampler is
    keyed by (``seed``, ``image_id``); ``image_id`` defaults to the number of images
    this process has pushed through the function, so that an unchanged
    run_inference.py gives every image its own random stream."""
    if image_id is None:
        image_id = _image_counter[0]
    _image_counter[0] += 1
    prediction_dict = model(sample_dict[IMAGE_NORMALIZED_KEY], train_val_test='testing')        # :22-23
    cls = prediction_dict[ANCHORS_CLASS_PREDICTIONS_KEY]                                          # [N,A,K]
    box = prediction_dict[ANCHORS_BOX_PREDICTIONS_KEY]                                            # [N,A,4]
    cov = prediction_dict.get(ANCHORS_COVAR_PREDICTIONS_KEY) if hasattr(prediction_dict, "get") else None   # :62
    anchors = sample_dict[ANCHORS_KEY]                                                            # [1,A,4]
    N, A, K = _shape(cls)

    cov_layout = _cabi.COV_NONE
    if cov is not None:
        cs = _shape(cov)
        cov_layout = _cabi.COV_FULL16 if cs[-2:] == (4, 4) else _cabi.COV_PACKED10
    sv = su = 1.0
    if dataset_name == 'kitti':                                                                   # :147-167
        orig = np.asarray(sample_dict[ORIGINAL_IM_SIZE_KEY]).reshape(-1)[:2].astype(np.float64)
        shp = np.asarray(_shape(sample_dict[IMAGE_NORMALIZED_KEY])[1:3], np.float64)
        sv, su = (float(np.float32(orig[0] / shp[0])), float(np.float32(orig[1] / shp[1])))
    cfg = BayesODConfig.from_reference(bayes_od_config, nms_config, use_full_covar, cov_layout=cov_layout,
                                       scale_v=sv, scale_u=su, seed=seed, image_id_base=image_id)
    eng = _engine(1, N, A, K, cfg, device)

    if _is_host(cls):                       # a host producer: stage through PCIe inside the library
        res = eng.run_host(_to_numpy(cls), _to_numpy(box), _to_numpy(cov), _to_numpy(anchors, (A, 4)), _to_numpy(counts))
    else:                                   # device tensors (TF via DLPack, torch, cupy): zero copy
        import torch                        # device memory plumbing only
        anc = anchors if not _is_host(anchors) else torch.as_tensor(_to_numpy(anchors, (A, 4))).cuda(device)
        cnt = None
        if counts is not None:
            cnt = counts if not _is_host(counts) else torch.as_tensor(_to_numpy(counts, (1, A, K))).cuda(device)
        eng.synchronize()                   # the producer's stream is not ours (TF does not expose it)
        eng.run(cls, box, cov, anc, cnt)
        res = eng.fetch()

    S, D = int(res.num_survivors[0]), int(res.num_dets[0])
    sv_ = eng.survivors(0) if S > 0 else None
    if S == 0:
        counts_out = np.zeros((0, K), np.float32); means = np.zeros((0, 4, 1), np.float32)
        covs = np.zeros((0, 4, 4), np.float32)
    else:
        counts_out, means, covs = sv_["counts"], sv_["means"][:, :, None], sv_["covs"]
    nms_indices = res.nms_indices[0, :D].astype(np.int32)
    scores, fmeans, fcovs, fcounts = res.image(0)
    token = FusedClusters(scores, fmeans, fcovs, fcounts, nms_indices, float(nms_config['iou_threshold']), S)
    return _ht(counts_out), _ht(means), _ht(covs), _ht(nms_indices), token


def bayes_od_clustering(predicted_boxes_class_counts, predicted_boxes_means, predicted_boxes_covs, cluster_centers,
                        affinity_matrix, affinity_threshold=0.7):
    """Bayesian NMS clustering (reference :285-364).  Returns float32 arrays
    (final_box_class_scores [D,K], final_box_means [D,4,1], final_box_covs [D,4,4]
    already x70, final_box_class_counts [D,K])."""
    if isinstance(affinity_matrix, FusedClusters):
        tok = affinity_matrix
        if not np.array_equal(np.asarray(cluster_centers).reshape(-1), tok.nms_indices):
            raise ValueError("cluster_centers differ from the nms_indices this FusedClusters token was built for")
        if abs(float(affinity_threshold) - tok.threshold) > 0:
            raise ValueError(f"affinity_threshold {affinity_threshold} differs from the nms_config['iou_threshold'] "
                             f"{tok.threshold} the clusters were fused with (run_inference.py:149 passes the same value)")
        return tok.scores, tok.means, tok.covs, tok.counts
    counts = np.ascontiguousarray(predicted_boxes_class_counts, np.float32)
    S, K = counts.shape
    centres = np.asarray(cluster_centers, np.int32).reshape(-1)
    cfg = BayesODConfig(max_output_size=255)
    cap = 1 << max(5, int(np.ceil(np.log2(max(S, 1)))))        # bucket the capacity: few distinct workspaces
    eng = _engine(1, 2, cap, K, cfg)
    return eng.cluster_host(counts, predicted_boxes_means, predicted_boxes_covs, centres, affinity_matrix,
                            float(affinity_threshold))
