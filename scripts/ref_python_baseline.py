"""Times the REFERENCE's own Python for the path -- inference_utils.bayes_od_inference (minus the model call) +
bayes_od_clustering, executed verbatim from /root/reference over the numpy-backed `tf` shim of
tests/golden/tf_numpy_shim.py -- on full-size synthetic images, as a second, labelled CPU baseline next to the C
oracle port (kind: "reference-python-over-numpy-shim").  Runs only where /root/reference exists (the build
container, not the GPU box); the result is committed under profiles/.  Real TensorFlow is not installable
offline (profiles/tf_probe_r2.txt), so the TF kernels themselves are numpy here: the number says what the
reference's algorithm costs in vectorised numpy on these host cores, not what TF's CPU kernels would reach."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import tf_numpy_shim as shim            # noqa: E402
import make_golden as mg                # noqa: E402
from bayes_od_rc_b200 import synthetic  # noqa: E402


def main():
    iu, bu, ag, cs, Categorical = shim.load_reference()
    out = {}
    for name, spec_kw in (("bdd_covar_k8", dict(N=10, K=8, config_id=3)), ("bdd_covar_k11", dict(N=10, K=11, config_id=3))):
        spec = synthetic.SceneSpec(**spec_kw)
        gen = ag.FpnAnchorGenerator(dict(aspect_ratios=[[1.0, 1.0], [1.0, 2.0], [2.0, 1.0]], scales=[1.0, 1.26, 1.59]))
        image_norm = np.zeros((spec.im_h, spec.im_w, 3), np.float32)
        anchors = np.concatenate([np.asarray(gen.generate_anchors(shim._t(np.asarray(image_norm.shape, np.int32)), l))
                                  for l in [3, 4, 5, 6, 7]], axis=0).astype(np.float32)
        cfg = json.loads(json.dumps(mg.BDD_TEST_CFG))
        times = []
        for image_id in range(3):
            img = synthetic.make_image(spec, image_id, torch.from_numpy(anchors), "cpu", with_counts=True)
            pred = {cs.ANCHORS_CLASS_PREDICTIONS_KEY: shim._t(img["cls"].numpy()), cs.ANCHORS_BOX_PREDICTIONS_KEY: shim._t(img["box"].numpy()),
                    cs.ANCHORS_COVAR_PREDICTIONS_KEY: shim._t(img["cov"].numpy())}
            model = lambda image, train_val_test='testing': pred   # noqa: E731
            sample_dict = {cs.IMAGE_NORMALIZED_KEY: shim._t(image_norm[None]), cs.ANCHORS_KEY: shim._t(anchors[None]),
                           cs.ORIGINAL_IM_SIZE_KEY: shim._t(np.asarray([[spec.im_h, spec.im_w, 3]], np.int32))}
            Categorical.forced_samples = mg.ids_from_counts(img["counts"].numpy().astype(np.float32), 30)
            t0 = time.perf_counter()
            o = iu.bayes_od_inference(model, sample_dict, cfg["bayes_od_config"], cfg["nms_config"], use_full_covar=True, dataset_name="bdd")
            cnt, mu, sig, idx, iou = [np.asarray(x) for x in o]
            iu.bayes_od_clustering(cnt, mu, sig, idx, iou, affinity_threshold=0.5)
            times.append(time.perf_counter() - t0)
            print(name, image_id, "S", len(cnt), "D", len(idx), f"{times[-1]:.2f} s", flush=True)
        out[name] = {"seconds_per_image": [round(t, 3) for t in times], "images_per_s": round(1.0 / (sum(times[1:]) / len(times[1:])), 3)}
    res = {"kind": "reference-python-over-numpy-shim", "unit": "images/s", "cores": os.cpu_count(),
           "where": "build container (no GPU); /root/reference is not present on the GPU box",
           "what": "src/retina_net/experiments/inference_utils.py bayes_od_inference (:25-217, model call replaced by a constant) + "
                   "bayes_od_clustering (:285-364), executed verbatim over tests/golden/tf_numpy_shim.py; 720x1280, N=10, full covariance; "
                   "first image is warm-up", "numpy": np.__version__, "workloads": out}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "ref_python_over_shim_r2.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
