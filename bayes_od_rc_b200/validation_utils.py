"""Drop-in for ``post_process_predictions`` of the reference's
``src/retina_net/experiments/validation_utils.py`` (:10-77), the deterministic
single-sample post-process that ``run_validation.py:143-147`` calls:

    post_process_predictions(sample_dict, prediction_dict, dataset_name='bdd')
        -> (predicted_boxes_classes_out [D,K], predicted_boxes_corners_out [D,4])

Same name, same arguments, same two return values (objects with ``.numpy()``),
so ``run_validation.py`` and its writers (``predictions_to_kitti_format``, the
BDD / COCO json builders) run unchanged.  Inside, the whole function — softmax,
arg-max filter, box decoding, soft-NMS, dataset-specific rescaling, gather — is
one ``bod_validate_run`` on the GPU; the prediction tensors are read in place
(DLPack / ``data_ptr``), nothing is computed on the host.

Install over the reference with::

    import bayes_od_rc_b200.validation_utils as fast
    from src.retina_net.experiments import validation_utils
    validation_utils.post_process_predictions = fast.post_process_predictions
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .engine import BayesODConfig
from .inference_utils import (ANCHORS_BOX_PREDICTIONS_KEY, ANCHORS_CLASS_PREDICTIONS_KEY, ANCHORS_KEY,
                              IMAGE_NORMALIZED_KEY, ORIGINAL_IM_SIZE_KEY, _engine, _ht, _is_host, _shape, _to_numpy)

IMAGE_PADDING_KEY = 'paddings_applied'              # src/core/constants.py:55


def _scaling(sample_dict, dataset_name):
    """validation_utils.py:54-66 as a bod_val_scaling tuple."""
    if dataset_name == 'kitti':
        shp = _shape(sample_dict[IMAGE_NORMALIZED_KEY])[1:3]                      # tf.shape(image[0])[:2]
        orig = np.asarray(sample_dict[ORIGINAL_IM_SIZE_KEY]).reshape(-1)[:2]
        return (_cabi.VAL_SCALE_KITTI, (0, 0, 0, 0), (float(shp[0]), float(shp[1])), (float(orig[0]), float(orig[1])))
    if dataset_name == 'coco':
        pad = _to_numpy(sample_dict[IMAGE_PADDING_KEY]).reshape(-1, 4)[0]
        shp = np.asarray(_shape(sample_dict[IMAGE_NORMALIZED_KEY])[1:3], np.int32) - (2 * pad[0:2]).astype(np.int32)
        orig = np.asarray(sample_dict[ORIGINAL_IM_SIZE_KEY]).reshape(-1)[:2]
        return (_cabi.VAL_SCALE_COCO, tuple(float(x) for x in pad), (float(shp[0]), float(shp[1])),
                (float(orig[0]), float(orig[1])))
    return None


def post_process_predictions(sample_dict, prediction_dict, dataset_name='bdd', device=0):
    """See module docstring.  ``prediction_dict`` holds the batch-1 tensors
    ``anchors_class_predictions [1,A,K]`` and ``anchors_box_predictions [1,A,4]``
    (device tensors or numpy arrays); ``sample_dict['anchors']`` is [1,A,4]."""
    cls = prediction_dict[ANCHORS_CLASS_PREDICTIONS_KEY]
    box = prediction_dict[ANCHORS_BOX_PREDICTIONS_KEY]
    _, A, K = _shape(cls)
    cfg = BayesODConfig(max_output_size=100, iou_threshold=0.5, soft_nms_sigma=0.5)     # literals of :47-52
    eng = _engine(1, 1, A, K, cfg, device)
    import torch                                    # device memory plumbing only

    def dev(x, n):
        """Device view of a producer tensor: CUDA tensors (torch, DLPack capsules / objects) pass through, host
        tensors (numpy, CPU tensors of any framework -- tf.data hands `anchors` over on the host) are staged."""
        if type(x).__name__ == "PyCapsule" or not _is_host(x) and (hasattr(x, "is_cuda") or hasattr(x, "__dlpack__")):
            return x
        return torch.as_tensor(_to_numpy(x).reshape(n)).cuda(device)
    eng.synchronize()                               # the producer's stream is not ours
    eng.validate(dev(cls, A * K), dev(box, A * 4), dev(sample_dict[ANCHORS_KEY], A * 4), _scaling(sample_dict, dataset_name))
    res = eng.fetch()
    D = int(res.num_dets[0])
    return _ht(res.cat_param[0, :D].copy()), _ht(res.means[0, :D].copy())
