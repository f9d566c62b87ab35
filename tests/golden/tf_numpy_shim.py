"""A numpy-backed stand-in for the subset of the TensorFlow / TFP API that the
reference's BayesOD path touches, so that the reference's OWN source files

    src/retina_net/experiments/inference_utils.py   (bayes_od_inference, bayes_od_clustering)
    src/retina_net/anchor_generator/box_utils.py
    src/retina_net/anchor_generator/fpn_anchor_generator.py

can be imported and EXECUTED in a container without TensorFlow.  Used only by
tests/golden/make_golden.py (fixture generation) and by the CPU tests that
cross-check the oracle against the live reference when /root/reference exists.
It is test infrastructure: nothing in the product imports it.

Semantics that matter and how they are kept:
  * TF converts the non-tensor operand of a binary op to the tensor's dtype
    (python floats, lists, numpy float64 scalars become float32).  `T` is an
    ndarray subclass whose __array_ufunc__ does exactly that, so float32 stays
    float32 everywhere, as in the reference's graph.
  * tf.linalg.inv / det -> numpy.linalg (LAPACK, float32 in -> float32 out).
  * tf.image.non_max_suppression_with_scores -> python restatement of TF's
    NonMaxSuppressionV5 CPU kernel (non_max_suppression_op.cc), binary32
    arithmetic, exp correctly rounded.
  * tfp.distributions.Categorical(probs).sample(n) is unseeded in the
    reference; here it draws from a numpy Generator that make_golden seeds, and
    the drawn class ids are recorded so the counts can be injected elsewhere.
"""
from __future__ import annotations

import heapq
import math
import sys
import types

import numpy as np

f32 = np.float32


class T(np.ndarray):
    """float32-preserving ndarray (see module docstring)."""
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        tdt = None
        for x in inputs:
            if isinstance(x, T) and np.issubdtype(x.dtype, np.floating):
                tdt = x.dtype if tdt is None else np.promote_types(tdt, x.dtype)
        conv = []
        for x in inputs:
            if isinstance(x, T):
                conv.append(x.view(np.ndarray))
            else:
                a = np.asarray(x)
                if tdt is not None and (np.issubdtype(a.dtype, np.floating) or np.issubdtype(a.dtype, np.integer)) \
                        and not isinstance(x, np.ndarray):
                    a = a.astype(tdt)           # python scalars / lists / numpy scalars follow the tensor
                elif tdt is not None and isinstance(x, np.ndarray) and x.ndim == 0:
                    a = a.astype(tdt)
                conv.append(a)
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, T) else o for o in out)
        res = getattr(ufunc, method)(*conv, **kwargs)
        if isinstance(res, tuple):
            return tuple(_t(r) for r in res)
        return _t(res)

    def numpy(self):
        return np.asarray(self)


def _t(x, dtype=None):
    if isinstance(x, (list, tuple)) and any(isinstance(e, np.ndarray) for e in x):
        x = np.stack([np.asarray(e) for e in x])
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64 and not isinstance(x, np.ndarray):
        a = a.astype(f32)                        # TF's default float is float32
    return a.view(T)


def _raw(x):
    return np.asarray(x)


def _ints(seq):
    return [int(round(float(np.asarray(s)))) for s in seq]


# ----------------------------------------------------------------------------
# NonMaxSuppressionV5 (soft-NMS) restatement
# ----------------------------------------------------------------------------
def _tf_iou(b, i, j):
    ymin_i = min(b[i, 0], b[i, 2]); xmin_i = min(b[i, 1], b[i, 3])
    ymax_i = max(b[i, 0], b[i, 2]); xmax_i = max(b[i, 1], b[i, 3])
    ymin_j = min(b[j, 0], b[j, 2]); xmin_j = min(b[j, 1], b[j, 3])
    ymax_j = max(b[j, 0], b[j, 2]); xmax_j = max(b[j, 1], b[j, 3])
    area_i = f32(ymax_i - ymin_i) * f32(xmax_i - xmin_i)
    area_j = f32(ymax_j - ymin_j) * f32(xmax_j - xmin_j)
    if area_i <= 0 or area_j <= 0:
        return f32(0)
    iymin = max(ymin_i, ymin_j); ixmin = max(xmin_i, xmin_j)
    iymax = min(ymax_i, ymax_j); ixmax = min(xmax_i, xmax_j)
    inter = f32(max(f32(iymax - iymin), f32(0))) * f32(max(f32(ixmax - ixmin), f32(0)))
    return f32(inter / f32(f32(area_i + area_j) - inter))


def non_max_suppression_with_scores(boxes, scores, max_output_size, iou_threshold=0.5,
                                    score_threshold=float("-inf"), soft_nms_sigma=0.0, name=None):
    b = _raw(boxes).astype(f32)
    s = _raw(scores).astype(f32)
    thr = f32(iou_threshold); sthr = f32(score_threshold); sigma = f32(soft_nms_sigma)
    heap = []
    for i in range(len(s)):
        if s[i] > sthr:
            heap.append((-float(s[i]), i, 0))   # max score first, ties -> lower index
    heapq.heapify(heap)
    is_soft = sigma > 0
    scale = f32(-0.5) / sigma if is_soft else f32(0)
    selected, sel_scores = [], []
    while len(selected) < int(max_output_size) and heap:
        neg, box, begin = heapq.heappop(heap)
        score = f32(-neg); original = score
        hard = False
        for j in range(len(selected) - 1, begin - 1, -1):
            sim = _tf_iou(b, box, selected[j])
            w = f32(math.exp(float(f32(f32(scale * sim) * sim))))
            if not (is_soft or sim <= thr):
                w = f32(0)
            score = f32(score * w)
            if (not is_soft) and sim > thr:
                hard = True
                break
            if score <= sthr:
                break
        begin = len(selected)
        if not hard:
            if score == original:
                selected.append(box); sel_scores.append(score)
                continue
            if score > sthr:
                heapq.heappush(heap, (-float(score), box, begin))
    return _t(np.asarray(selected, np.int32)), _t(np.asarray(sel_scores, f32))


# ----------------------------------------------------------------------------
# module assembly
# ----------------------------------------------------------------------------
class _Categorical:
    rng = np.random.default_rng(0)
    last_samples = None
    forced_samples = None        # [n, A] class ids to return instead of drawing

    def __init__(self, probs=None, logits=None):
        self.probs = _raw(probs).astype(np.float64)

    def sample(self, n):
        if _Categorical.forced_samples is not None:
            ids = np.asarray(_Categorical.forced_samples, np.int32)
            assert ids.shape == (int(n), self.probs.shape[0])
            _Categorical.last_samples = ids
            return _t(ids)
        p = self.probs / self.probs.sum(axis=-1, keepdims=True)
        cdf = np.cumsum(p, axis=-1)
        u = _Categorical.rng.random((int(n), p.shape[0], 1))
        ids = (u >= cdf[None]).sum(axis=-1)
        ids = np.minimum(ids, p.shape[-1] - 1).astype(np.int32)
        _Categorical.last_samples = ids
        return _t(ids)


def fill_triangular(x):
    """tfp.math.fill_triangular (lower): rows of concat(x[n:], reverse(x)) reshaped n x n."""
    x = _raw(x)
    m = x.shape[-1]
    n = int(round((math.sqrt(8 * m + 1) - 1) / 2))
    full = np.concatenate([x[..., n:], x[..., ::-1]], axis=-1).reshape(x.shape[:-1] + (n, n))
    return _t(np.tril(full))


def _reduce_sum(x, axis=None, keepdims=False):
    """tf.reduce_sum.  The order in which TensorFlow adds the terms is internal to its (Eigen / GPU) reduction kernels and
    differs between devices and versions; for binary32 it matters as soon as the terms are inexact, e.g. the Dirichlet
    counts + 1/K at inference_utils.py:96-97 with K = 11, whose row sums feed the ranking score and through it the order of
    soft-NMS selections among candidates one ulp apart.  This build fixes the order the way its arithmetic contract does
    (reductions sequential in index order), and the shim follows it for reductions over one axis of a float
    array, so that the fixtures pin everything else; numpy's own pairwise order would be just as arbitrary a stand-in."""
    a = _raw(x)
    if axis is None or not np.issubdtype(a.dtype, np.floating) or isinstance(axis, (tuple, list)):
        return _t(np.sum(a, axis=axis, keepdims=keepdims))
    m = np.moveaxis(a, axis, 0)
    acc = m[0].copy()
    for k in range(1, m.shape[0]):
        acc = acc + m[k]                      # one correctly rounded addition per term, in index order
    if keepdims:
        acc = np.expand_dims(acc, axis)
    return _t(acc)


def install():
    """Put fake `tensorflow` / `tensorflow_probability` modules in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = np.float32, np.float64, np.int32, np.int64, np.bool_
    tf.function = lambda fn=None, **kw: fn if fn is not None else (lambda f: f)
    tf.constant = lambda v, dtype=None: _t(v, dtype)
    tf.convert_to_tensor = tf.constant
    tf.cast = lambda x, dt: _t(_raw(x).astype(dt))
    tf.shape = lambda x: _t(np.asarray(_raw(x).shape, np.int32))
    tf.size = lambda x: _t(np.asarray(_raw(x).size, np.int32))
    tf.equal = lambda a, b: _t(np.equal(_raw(a), _raw(b)))
    tf.not_equal = lambda a, b: _t(np.not_equal(_raw(a), _raw(b)))
    tf.greater_equal = lambda a, b: _t(np.greater_equal(_raw(a), _raw(b)))
    tf.less_equal = lambda a, b: _t(np.less_equal(_raw(a), _raw(b)))
    tf.argmax = lambda x, axis=None, name=None: _t(np.argmax(_raw(x), axis=axis).astype(np.int64))
    tf.reduce_mean = lambda x, axis=None, keepdims=False: _t(np.mean(_raw(x), axis=axis, keepdims=keepdims, dtype=_raw(x).dtype))
    tf.reduce_sum = _reduce_sum
    tf.reduce_max = lambda x, axis=None, keepdims=False: _t(np.max(_raw(x), axis=axis, keepdims=keepdims))
    tf.reduce_min = lambda x, axis=None, keepdims=False: _t(np.min(_raw(x), axis=axis, keepdims=keepdims))
    tf.reduce_any = lambda x, axis=None: _t(np.any(_raw(x), axis=axis))
    tf.reduce_all = lambda x, axis=None: _t(np.all(_raw(x), axis=axis))
    tf.exp = lambda x: _t(np.exp(_t(x)))
    tf.sqrt = lambda x: _t(np.sqrt(_t(x)))
    tf.pow = lambda a, b: _t(np.power(_t(a), _t(b)))
    tf.maximum = lambda a, b: _t(np.maximum(_t(a), _t(b) if not isinstance(b, T) else b))
    tf.minimum = lambda a, b: _t(np.minimum(_t(a), _t(b) if not isinstance(b, T) else b))
    tf.clip_by_value = lambda x, lo, hi: _t(np.minimum(np.maximum(_t(x), f32(lo)), f32(hi)))
    tf.zeros_like = lambda x: _t(np.zeros_like(_raw(x)))
    tf.ones_like = lambda x: _t(np.ones_like(_raw(x)))
    tf.ones = lambda shape, dtype=np.float32: _t(np.ones(_ints(shape), dtype))
    tf.zeros = lambda shape, dtype=np.float32: _t(np.zeros(_ints(shape), dtype))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(_t(x), axis))
    tf.squeeze = lambda x, axis=None: _t(np.squeeze(_raw(x), axis=axis))
    tf.transpose = lambda x, perm=None: _t(np.transpose(_raw(x), perm))
    tf.reshape = lambda x, shape: _t(np.reshape(_raw(x), _ints(shape)))
    tf.stack = lambda xs, axis=0: _t(np.stack([_raw(_t(x)) for x in xs], axis=axis))
    tf.concat = lambda xs, axis=0: _t(np.concatenate([_raw(x) for x in xs], axis=axis))
    tf.split = lambda x, n, axis=0: [_t(p) for p in np.split(_raw(x), n, axis=axis)]
    tf.tile = lambda x, multiples: _t(np.tile(_raw(_t(x)), _ints(multiples)))
    tf.range = lambda start, limit=None, delta=1: _t(
        np.arange(float(_raw(start)), float(_raw(limit)), float(delta)).astype(f32)
        if np.issubdtype(_raw(limit).dtype, np.floating) else np.arange(int(start), int(limit), int(delta)))
    tf.meshgrid = lambda *xs: [_t(m) for m in np.meshgrid(*[_raw(x) for x in xs])]
    tf.gather = lambda x, idx, axis=0: _t(np.take(_raw(x), _raw(idx), axis=axis))
    tf.where = lambda c, a, b: _t(np.where(_raw(c), _raw(a), _raw(b)))
    tf.broadcast_to = lambda x, shape: _t(np.broadcast_to(_raw(x), _ints(shape)))

    def boolean_mask(tensor, mask, axis=0):
        t, m = _raw(tensor), _raw(mask).astype(bool)
        return _t(np.compress(m, t, axis=axis))
    tf.boolean_mask = boolean_mask

    def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=None, dtype=None):
        idx = _raw(indices)
        out = np.full(idx.shape + (int(depth),), f32(off_value), f32)
        np.put_along_axis(out, idx[..., None].astype(np.int64), f32(on_value), axis=-1)
        return _t(out)
    tf.one_hot = one_hot

    def matmul(a, b, transpose_a=False, transpose_b=False):
        a, b = _raw(_t(a)), _raw(_t(b))
        if transpose_a:
            a = np.swapaxes(a, -1, -2)
        if transpose_b:
            b = np.swapaxes(b, -1, -2)
        return _t(np.matmul(a, b))
    tf.matmul = matmul

    class _Scope:
        def __init__(self, *a, **k): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    tf.name_scope = _Scope

    nn = types.ModuleType("tensorflow.nn")

    def softmax(x, axis=-1):
        x = _raw(x)
        sh = x - np.max(x, axis=axis, keepdims=True)
        e = np.exp(sh)
        return _t(e / np.sum(e, axis=axis, keepdims=True))
    nn.softmax = softmax
    tf.nn = nn

    mth = types.ModuleType("tensorflow.math")
    mth.log = lambda x: _t(np.log(_t(x)))
    mth.exp = tf.exp
    tf.math = mth

    la = types.ModuleType("tensorflow.linalg")
    la.inv = lambda x: _t(np.linalg.inv(_raw(x)))
    la.det = lambda x: _t(np.linalg.det(_raw(x)).astype(_raw(x).dtype))
    la.diag_part = lambda x: _t(np.diagonal(_raw(x), axis1=-2, axis2=-1).copy())

    def set_diag(x, diag):
        out = _raw(x).copy()
        n = out.shape[-1]
        out[..., np.arange(n), np.arange(n)] = _raw(diag)
        return _t(out)
    la.set_diag = set_diag
    la.tensor_diag = lambda d: _t(np.diag(_raw(_t(d))))
    tf.linalg = la

    img = types.ModuleType("tensorflow.image")
    img.non_max_suppression_with_scores = non_max_suppression_with_scores
    tf.image = img

    tfp = types.ModuleType("tensorflow_probability")
    dist = types.ModuleType("tensorflow_probability.distributions")
    dist.Categorical = _Categorical
    tfp.distributions = dist
    tmath = types.ModuleType("tensorflow_probability.math")
    tmath.fill_triangular = fill_triangular
    tfp.math = tmath

    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow_probability"] = tfp
    return tf, tfp


def load_reference(ref_root="/root/reference"):
    """Import the reference's own modules over the shim. Returns
    (inference_utils, box_utils, fpn_anchor_generator, constants, Categorical)."""
    install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import importlib
    iu = importlib.import_module("src.retina_net.experiments.inference_utils")
    bu = importlib.import_module("src.retina_net.anchor_generator.box_utils")
    ag = importlib.import_module("src.retina_net.anchor_generator.fpn_anchor_generator")
    cs = importlib.import_module("src.core.constants")
    return iu, bu, ag, cs, _Categorical
