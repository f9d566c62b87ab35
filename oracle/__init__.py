"""ctypes binding of the CPU oracle (oracle/bayesod_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, from bench.py's cpu_baseline /
``--impl reference`` arms and from ``__graft_entry__.smoke()``; the product
package (bayes_od_rc_b200) never imports it.

Every wrapper runs ONE image (the reference is batch-1,
run_inference.py:68); ``run_image`` chains them exactly as
inference_utils.py:25-217 + :285-364 do and returns every intermediate.
``real='f64'`` selects the binary64 twin (adjudication of ill-conditioned
elements, SURVEY.md §7 hard part 3).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbayesod_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle). Returns the .so path."""
    stale = False
    for name, out in (("bayesod_oracle.c", _LIB_PATH), ("pdq_oracle.c", os.path.join(_HERE, "_build", "libpdq_oracle.so"))):
        src = os.path.join(_HERE, name)
        stale = stale or (not os.path.exists(out)) or os.path.getmtime(out) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [("N", C.c_int32), ("A", C.c_int32), ("K", C.c_int32),
                ("cov_layout", C.c_int32), ("use_full_covar", C.c_int32),
                ("dirichlet_prior", C.c_int32), ("gaussian_prior", C.c_int32),
                ("isotropic_variance", C.c_float), ("ranking_method", C.c_int32),
                ("scale_v", C.c_float), ("scale_u", C.c_float)]


class _RunParams(C.Structure):
    _fields_ = [("p", _Params), ("B", C.c_int32), ("max_output_size", C.c_int32),
                ("iou_threshold", C.c_float), ("soft_nms_sigma", C.c_float),
                ("cov_calibration", C.c_float), ("num_draws", C.c_int32),
                ("seed", C.c_uint64), ("image_id_base", C.c_uint32),
                ("score_threshold", C.c_float), ("pre_nms_top_k", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc32_det4.restype = C.c_float
        _lib.orc64_det4.restype = C.c_double
        _lib.orc32_tf_iou.restype = C.c_float
        _lib.orc64_tf_iou.restype = C.c_double
        _lib.orc32_repo_iou.restype = C.c_float
        _lib.orc64_repo_iou.restype = C.c_double
    return _lib


@dataclass
class OracleConfig:
    """The knobs run_inference.py:25-29 reads from testing_config (defaults =
    retinanet_bdd_covar.yaml:117-144) plus the extension knobs of SURVEY §8(d)."""
    use_full_covar: bool = True
    cov_layout: int = 1                 # 0 none, 1 [N,A,4,4], 2 packed [N,A,10]
    dirichlet_prior: str = "non_informative"
    gaussian_prior: str = "isotropic"
    isotropic_variance: float = 100000.0
    ranking_method: str = "score"
    max_output_size: int = 100
    iou_threshold: float = 0.5
    soft_nms_sigma: float = 0.5
    scale_v: float = 1.0
    scale_u: float = 1.0
    cov_calibration: float = 70.0
    num_draws: int = 30
    seed: int = 1234
    image_id_base: int = 0
    score_threshold: float = -np.inf
    pre_nms_top_k: int = 0

    def params(self, N, A, K) -> _Params:
        return _Params(N, A, K, self.cov_layout, int(self.use_full_covar),
                       1 if self.dirichlet_prior == "non_informative" else 0,
                       1 if self.gaussian_prior == "isotropic" else 0,
                       float(self.isotropic_variance),
                       1 if (self.ranking_method == "joint_entropy" and self.gaussian_prior != "None"
                             and self.dirichlet_prior != "None") else 0,
                       float(self.scale_v), float(self.scale_u))


def _f(a, dtype=np.float32):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _rt(real):
    return (np.float64, "orc64_") if real == "f64" else (np.float32, "orc32_")


# --------------------------------------------------------------------------
# stage wrappers
# --------------------------------------------------------------------------
def generate_anchors(im_h: int, im_w: int) -> np.ndarray:
    """fpn_anchor_generator.py:21-59, levels 3..7 concatenated P3->P7."""
    L = lib()
    A = L.orc_generate_anchors(int(im_h), int(im_w), None)
    out = np.empty((A, 4), np.float32)
    L.orc_generate_anchors(int(im_h), int(im_w), _p(out))
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def softmax_mean(cls: np.ndarray, real="f32") -> np.ndarray:
    """inference_utils.py:31-32,38. cls [N,A,K] -> [A,K]."""
    dt, pre = _rt(real)
    cls = _f(cls); N, A, K = cls.shape
    out = np.empty((A, K), dt)
    getattr(lib(), pre + "softmax_mean")(_p(cls), N, A, K, _p(out))
    return out


def philox_counts(probs: np.ndarray, num_draws=30, seed=1234, image_id=0) -> np.ndarray:
    """The product's documented Philox sampler, restated (stands in for :37-46)."""
    probs = _f(probs); A, K = probs.shape
    out = np.empty((A, K), np.float32)
    lib().orc_philox_counts(_p(probs), A, K, int(num_draws), C.c_uint64(int(seed)), C.c_uint32(int(image_id)), _p(out))
    return out


def category_filter(counts: np.ndarray) -> np.ndarray:
    """inference_utils.py:48-54 -> ascending kept anchor indices."""
    counts = _f(counts); A, K = counts.shape
    keep = np.empty(A, np.int32)
    S = lib().orc_filter(_p(counts), A, K, _p(keep))
    return keep[:S].copy()


def prefilter(counts, keep, K, dirichlet=True, score_threshold=-np.inf, top_k=0) -> np.ndarray:
    counts = _f(counts); keep = np.ascontiguousarray(keep, np.int32).copy()
    S = lib().orc_prefilter(_p(counts), K, int(dirichlet), C.c_float(score_threshold), int(top_k), _p(keep), len(keep))
    return keep[:S].copy()


def posterior(cfg: OracleConfig, box, cov, anchors, counts, keep, real="f32"):
    """inference_utils.py:28-29, 57-205 for the kept anchors."""
    dt, pre = _rt(real)
    box = _f(box); N, A, _ = box.shape
    counts = _f(counts); K = counts.shape[1]
    cov_c = None
    if cfg.cov_layout != 0:
        cov_c = _f(cov).reshape(N, A, 16 if cfg.cov_layout == 1 else 10)
    anchors = _f(anchors).reshape(A, 4)
    keep = np.ascontiguousarray(keep, np.int32); S = len(keep)
    P = cfg.params(N, A, K)
    cp = np.empty((S, K), dt); mu = np.empty((S, 4), dt); sg = np.empty((S, 16), dt)
    sc = np.empty((S,), dt); co = np.empty((S, 4), dt)
    getattr(lib(), pre + "posterior")(C.byref(P), _p(box), _p(cov_c), _p(anchors), _p(counts), _p(keep), S,
                                      _p(cp), _p(mu), _p(sg), _p(sc), _p(co))
    return cp, mu, sg.reshape(S, 4, 4), sc, co


def nms_v5(corners, scores, max_output_size=100, iou_threshold=0.5, score_threshold=-np.inf,
           soft_nms_sigma=0.5, real="f32", return_stats=False):
    """tf.image.non_max_suppression_with_scores (NonMaxSuppressionV5)."""
    dt, pre = _rt(real)
    corners = _f(corners, dt).reshape(-1, 4); scores = _f(scores, dt); S = len(scores)
    sel = np.empty(max(max_output_size, 1), np.int32); ss = np.empty(max(max_output_size, 1), dt)
    stats = np.zeros(2, np.int64)
    D = getattr(lib(), pre + "nms_v5")(_p(corners), _p(scores), S, int(max_output_size), C.c_float(iou_threshold),
                                       C.c_float(score_threshold), C.c_float(soft_nms_sigma), _p(sel), _p(ss), _p(stats))
    if return_stats:
        return sel[:D].copy(), ss[:D].copy(), stats
    return sel[:D].copy(), ss[:D].copy()


def iou_matrix(corners, real="f32") -> np.ndarray:
    """box_utils.bbox_iou_vuvu(c, c) (box_utils.py:117-146), full [S,S]."""
    dt, pre = _rt(real)
    corners = _f(corners, dt).reshape(-1, 4); S = len(corners)
    out = np.empty((S, S), dt)
    getattr(lib(), pre + "iou_matrix")(_p(corners), S, _p(out))
    return out


def membership(corners, centres, thr=0.5, real="f32") -> np.ndarray:
    """Bit s of row d <=> affinity[s, centre_d] > thr (inference_utils.py:316)."""
    dt, pre = _rt(real)
    corners = _f(corners, dt).reshape(-1, 4); S = len(corners)
    centres = np.ascontiguousarray(centres, np.int32); D = len(centres)
    wpr = max((S + 31) // 32, 1)
    mask = np.zeros((max(D, 1), wpr), np.uint32)
    getattr(lib(), pre + "membership")(_p(corners), S, _p(centres), D, C.c_float(thr), _p(mask), wpr)
    return mask[:D]


def mask_to_bool(mask: np.ndarray, S: int) -> np.ndarray:
    bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :S].astype(bool)


def clustering(cnt, mu, sig, centres, mask, calibration=70.0, real="f32"):
    """bayes_od_clustering (inference_utils.py:285-364) with membership given as bitmask."""
    dt, pre = _rt(real)
    cnt = _f(cnt, dt); S, K = cnt.shape
    mu = _f(mu, dt).reshape(S, 4); sig = _f(sig, dt).reshape(S, 16)
    centres = np.ascontiguousarray(centres, np.int32); D = len(centres)
    mask = np.ascontiguousarray(mask, np.uint32).reshape(max(D, 1), -1) if D else np.zeros((1, 1), np.uint32)
    D1 = max(D, 1)
    os_ = np.zeros((D1, K), dt); om = np.zeros((D1, 4), dt); oc = np.zeros((D1, 16), dt); on = np.zeros((D1, K), dt)
    mem = np.zeros(D1, np.int32); chosen = np.full((D1, 3), -1, np.int32)
    empty = getattr(lib(), pre + "clustering")(_p(cnt), _p(mu), _p(sig), S, K, _p(centres), D, _p(mask), mask.shape[1],
                                               C.c_float(calibration), _p(os_), _p(om), _p(oc), _p(on), _p(mem), _p(chosen))
    clustering.last_chosen = chosen[:D]
    return os_[:D], om[:D], oc[:D].reshape(D, 4, 4), on[:D], mem[:D], empty


# --------------------------------------------------------------------------
# the whole path for one image, every intermediate kept
# --------------------------------------------------------------------------
@dataclass
class ImageResult:
    probs: np.ndarray = None            # [A,K]   mean class probabilities
    counts: np.ndarray = None           # [A,K]   categorical sample counts (injected or Philox)
    keep: np.ndarray = None             # [S]     kept anchor indices
    cnt_post: np.ndarray = None         # [S,K]
    mu_post: np.ndarray = None          # [S,4]
    sig_post: np.ndarray = None         # [S,4,4]
    score: np.ndarray = None            # [S]
    corners: np.ndarray = None          # [S,4]
    nms_indices: np.ndarray = None      # [D]
    nms_scores: np.ndarray = None       # [D]
    mask: np.ndarray = None             # [D, ceil(S/32)] uint32
    members: np.ndarray = None          # [D]
    final_scores: np.ndarray = None     # [D,K]
    final_means: np.ndarray = None      # [D,4]
    final_covs: np.ndarray = None       # [D,4,4]
    final_counts: np.ndarray = None     # [D,K]
    empty_clusters: int = 0
    extra: dict = field(default_factory=dict)


def run_image(cfg: OracleConfig, cls, box, cov, anchors, counts=None, image_id=0, real="f32",
              with_probs=True, force=None) -> ImageResult:
    """cls [N,A,K], box [N,A,4], cov [N,A,4,4]|[N,A,10]|None, anchors [A,4],
    counts [A,K] or None (-> Philox sampler on the oracle's own probabilities).

    ``force`` (binary64 twin): dict with 'nms_indices' and 'mask' taken from the
    binary32 run so both precisions fuse the same clusters."""
    r = ImageResult()
    cls = _f(cls); N, A, K = cls.shape
    if with_probs or counts is None:
        r.probs = softmax_mean(cls, real)
    if counts is None:
        counts = philox_counts(r.probs.astype(np.float32), cfg.num_draws, cfg.seed, cfg.image_id_base + image_id)
    r.counts = _f(counts)
    r.keep = category_filter(r.counts)
    if cfg.pre_nms_top_k > 0 or cfg.score_threshold > -np.inf:
        r.keep = prefilter(r.counts, r.keep, K, cfg.dirichlet_prior == "non_informative",
                           cfg.score_threshold, cfg.pre_nms_top_k)
    r.cnt_post, r.mu_post, r.sig_post, r.score, r.corners = posterior(cfg, box, cov, anchors, r.counts, r.keep, real)
    if force is not None:
        r.nms_indices = np.asarray(force["nms_indices"], np.int32); r.mask = force["mask"]
        r.nms_scores = None
    else:
        r.nms_indices, r.nms_scores = nms_v5(r.corners, r.score, cfg.max_output_size, cfg.iou_threshold,
                                             -np.inf, cfg.soft_nms_sigma, real)
        r.mask = membership(r.corners, r.nms_indices, cfg.iou_threshold, real)
    (r.final_scores, r.final_means, r.final_covs, r.final_counts, r.members,
     r.empty_clusters) = clustering(r.cnt_post, r.mu_post, r.sig_post, r.nms_indices, r.mask, cfg.cov_calibration, real)
    r.extra["chosen"] = clustering.last_chosen
    return r


class _ValScaling(C.Structure):
    _fields_ = [("mode", C.c_int32), ("shift", C.c_float * 4), ("norm_h", C.c_float), ("norm_w", C.c_float),
                ("scale_h", C.c_float), ("scale_w", C.c_float)]


@dataclass
class ValResult:
    keep: np.ndarray = None        # [S] kept anchor indices
    probs: np.ndarray = None       # [S,K] softmax probabilities
    corners: np.ndarray = None     # [S,4] unscaled corners (what NMS sees)
    scores: np.ndarray = None      # [S]
    nms_indices: np.ndarray = None  # [D]
    nms_scores: np.ndarray = None
    classes_out: np.ndarray = None  # [D,K]
    corners_out: np.ndarray = None  # [D,4]


def val_postprocess(cls, box, anchors, max_output_size=100, iou_threshold=0.5, soft_nms_sigma=0.5,
                    scale_mode=0, shift=(0, 0, 0, 0), norm_hw=(1, 1), scale_hw=(1, 1)) -> ValResult:
    """validation_utils.post_process_predictions (validation_utils.py:10-77) for one image.
    cls [A,K] logits, box [A,4] deltas, anchors [A,4].  scale_mode 0 none / 1 kitti / 2 coco."""
    cls = _f(cls); A, K = cls.shape
    box = _f(box).reshape(A, 4); anchors = _f(anchors).reshape(A, 4)
    sc = _ValScaling(int(scale_mode), (C.c_float * 4)(*[float(x) for x in shift]), float(norm_hw[0]), float(norm_hw[1]),
                     float(scale_hw[0]), float(scale_hw[1]))
    D = int(max_output_size)
    keep = np.empty(A, np.int32); probs = np.empty((A, K), np.float32); corners = np.empty((A, 4), np.float32)
    scores = np.empty(A, np.float32); sel = np.empty(max(D, 1), np.int32); ss = np.empty(max(D, 1), np.float32)
    oc = np.zeros((max(D, 1), K), np.float32); ob = np.zeros((max(D, 1), 4), np.float32)
    nd = C.c_int32(0)
    S = lib().orc_val_postprocess(_p(cls), _p(box), _p(anchors), A, K, D, C.c_float(iou_threshold),
                                  C.c_float(soft_nms_sigma), C.byref(sc), _p(keep), _p(probs), _p(corners), _p(scores),
                                  _p(sel), _p(ss), _p(oc), _p(ob), C.byref(nd))
    d = nd.value
    return ValResult(keep[:S].copy(), probs[:S].copy(), corners[:S].copy(), scores[:S].copy(), sel[:d].copy(),
                     ss[:d].copy(), oc[:d].copy(), ob[:d].copy())


def run_batch(cfg: OracleConfig, cls, box, cov, anchors, counts=None, nthreads=1):
    """Whole path for [B,...] inputs, padded outputs (CPU baseline arm)."""
    cls = _f(cls); B, N, A, K = cls.shape
    box = _f(box)
    cov_c = None if cfg.cov_layout == 0 else _f(cov).reshape(B, N, A, -1)
    anchors = _f(anchors).reshape(A, 4)
    cnt = None if counts is None else _f(counts)
    D = cfg.max_output_size
    rp = _RunParams(cfg.params(N, A, K), B, D, cfg.iou_threshold, cfg.soft_nms_sigma, cfg.cov_calibration,
                    cfg.num_draws, cfg.seed, cfg.image_id_base, cfg.score_threshold, cfg.pre_nms_top_k)
    out = dict(num_dets=np.zeros(B, np.int32), num_survivors=np.zeros(B, np.int32),
               means=np.zeros((B, D, 4), np.float32), covs=np.zeros((B, D, 16), np.float32),
               cat_param=np.zeros((B, D, K), np.float32), cat_count=np.zeros((B, D, K), np.float32),
               nms_indices=np.zeros((B, D), np.int32), centre_anchor_idx=np.zeros((B, D), np.int32))
    lib().orc_run_batch(C.byref(rp), _p(cls), _p(box), _p(cov_c), _p(anchors), _p(cnt),
                        _p(out["num_dets"]), _p(out["num_survivors"]), _p(out["means"]), _p(out["covs"]),
                        _p(out["cat_param"]), _p(out["cat_count"]), _p(out["nms_indices"]),
                        _p(out["centre_anchor_idx"]), int(nthreads))
    return out
