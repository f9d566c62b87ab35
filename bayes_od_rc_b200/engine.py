"""Host-side driver of the C ABI: owns a ``bod_ctx`` and moves arrays in and out.

``BayesODEngine`` is what the drop-in functions of
``bayes_od_rc_b200.inference_utils`` and ``bench.py`` call.  Inputs may be
 * device tensors (anything with ``data_ptr()`` / ``__cuda_array_interface__`` /
   ``__dlpack__`` or a DLPack capsule, e.g. ``tf.experimental.dlpack.to_dlpack(t)``)
   -> ``run()`` (zero copy), or
 * numpy arrays -> ``run_host()`` (the library stages them through PCIe).
PyTorch is not imported here; it is only ever the owner of device memory in
tests and the bench.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from ._cabi import BodConfig, BodError, BodHostResults, BodHostSurvivors


@dataclass
class BayesODConfig:
    """testing_config of the reference's YAML (retinanet_bdd.yaml:117-144) plus the
    extension knobs; ``from_reference`` takes the very dicts run_inference.py:25-29 reads."""
    use_full_covar: bool = False
    cov_layout: int = _cabi.COV_FULL16
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
    score_threshold: float = float("-inf")
    pre_nms_top_k: int = 0
    anchor_mode: int = _cabi.ANCHORS_TENSOR
    im_h: int = 0
    im_w: int = 0
    max_survivors: int = 0
    emit_probs: bool = False
    pipeline_depth: int = 1
    level_anchors: tuple = ()       # anchors per FPN level (P3->P7) when the head outputs come per level (run_levels)

    @classmethod
    def from_reference(cls, bayes_od_config: dict, nms_config: dict, use_full_covar: bool = False, **kw):
        g = bayes_od_config["gaussian_prior"]
        return cls(use_full_covar=bool(use_full_covar),
                   dirichlet_prior=str(bayes_od_config["dirichlet_prior"]["type"]),
                   gaussian_prior=str(g["type"]),
                   isotropic_variance=float(g.get("isotropic_variance", 100000.0)),
                   ranking_method=str(bayes_od_config["ranking_method"]),
                   max_output_size=int(nms_config["max_output_size"]),
                   iou_threshold=float(nms_config["iou_threshold"]),
                   soft_nms_sigma=float(nms_config["soft_nms_sigma"]), **kw)

    def to_c(self, B, N, A, K) -> BodConfig:
        return BodConfig(
            B=B, N=N, A=A, K=K, cov_layout=self.cov_layout, use_full_covar=int(self.use_full_covar),
            dirichlet_prior=1 if self.dirichlet_prior == "non_informative" else 0,
            gaussian_prior=1 if self.gaussian_prior == "isotropic" else 0,
            isotropic_variance=self.isotropic_variance,
            ranking_method=1 if self.ranking_method == "joint_entropy" else 0,
            max_output_size=self.max_output_size, iou_threshold=self.iou_threshold,
            soft_nms_sigma=self.soft_nms_sigma, scale_v=self.scale_v, scale_u=self.scale_u,
            cov_calibration=self.cov_calibration, num_draws=self.num_draws, seed=self.seed,
            image_id_base=self.image_id_base, score_threshold=self.score_threshold,
            pre_nms_top_k=self.pre_nms_top_k, anchor_mode=self.anchor_mode, im_h=self.im_h, im_w=self.im_w,
            max_survivors=self.max_survivors, emit_probs=int(self.emit_probs),
            pipeline_depth=int(self.pipeline_depth), n_levels=len(self.level_anchors),
            level_anchors=(C.c_int32 * 8)(*[int(x) for x in self.level_anchors]))


# --------------------------------------------------------------------------
# device pointer extraction (torch / __cuda_array_interface__ / DLPack)
# --------------------------------------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


_kDLCUDA, _kDLFloat = 2, 2


def _from_dlpack_capsule(cap, expect_elems):
    api = C.pythonapi
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    ptr = api.PyCapsule_GetPointer(cap, b"dltensor")
    t = C.cast(ptr, C.POINTER(_DLTensor)).contents
    if t.device.device_type != _kDLCUDA:
        raise TypeError("DLPack tensor is not in CUDA device memory")
    if not (t.dtype.code == _kDLFloat and t.dtype.bits == 32 and t.dtype.lanes == 1):
        raise TypeError("DLPack tensor is not float32")
    shape = [t.shape[i] for i in range(t.ndim)]
    n = int(np.prod(shape)) if shape else 1
    if t.strides:                                   # must be C-contiguous
        expect = 1
        for i in range(t.ndim - 1, -1, -1):
            if shape[i] != 1 and t.strides[i] != expect:
                raise TypeError("DLPack tensor is not contiguous")
            expect *= shape[i]
    if expect_elems is not None and n != expect_elems:
        raise ValueError(f"tensor has {n} elements, expected {expect_elems}")
    return (t.data or 0) + t.byte_offset


def device_ptr(x, expect_elems=None, keepalive=None):
    """Address of a contiguous float32 CUDA array, without copying."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr") and hasattr(x, "is_cuda"):              # torch.Tensor
        if not x.is_cuda:
            raise TypeError("expected a CUDA tensor (use run_host for host arrays)")
        if str(x.dtype) != "torch.float32" or not x.is_contiguous():
            raise TypeError("expected a contiguous float32 tensor")
        if expect_elems is not None and x.numel() != expect_elems:
            raise ValueError(f"tensor has {x.numel()} elements, expected {expect_elems}")
        return x.data_ptr()
    if hasattr(x, "__cuda_array_interface__"):
        cai = x.__cuda_array_interface__
        if cai["typestr"] not in ("<f4", "=f4") or cai.get("strides") is not None:
            raise TypeError("expected a contiguous float32 CUDA array")
        if expect_elems is not None and int(np.prod(cai["shape"])) != expect_elems:
            raise ValueError("unexpected number of elements")
        return cai["data"][0]
    if type(x).__name__ == "PyCapsule":
        return _from_dlpack_capsule(x, expect_elems)
    if hasattr(x, "__dlpack__"):
        if hasattr(x, "__dlpack_device__") and int(x.__dlpack_device__()[0]) not in (2, 13):     # kDLCUDA, kDLCUDAManaged
            raise TypeError("DLPack tensor is not in CUDA device memory (stage host tensors through run_host)")
        cap = x.__dlpack__()
        if keepalive is not None:
            keepalive.append(cap)
        return _from_dlpack_capsule(cap, expect_elems)
    raise TypeError(f"cannot take a device pointer from {type(x)}")


# --------------------------------------------------------------------------
@dataclass
class Results:
    """Padded result blocks of a batch (bod_host_results)."""
    num_dets: np.ndarray
    num_survivors: np.ndarray
    means: np.ndarray          # [B,Dmax,4]
    covs: np.ndarray           # [B,Dmax,4,4]
    cat_param: np.ndarray      # [B,Dmax,K]
    cat_count: np.ndarray      # [B,Dmax,K]
    nms_indices: np.ndarray    # [B,Dmax]
    centre_anchor_idx: np.ndarray
    centre_scores: np.ndarray

    def image(self, b: int):
        """(class_scores[D,K], means[D,4,1], covs[D,4,4], class_counts[D,K]) as
        bayes_od_clustering returns them (inference_utils.py:364)."""
        d = int(self.num_dets[b])
        return (self.cat_param[b, :d].copy(), self.means[b, :d, :, None].copy(), self.covs[b, :d].copy(),
                self.cat_count[b, :d].copy())


class BayesODEngine:
    """One ``bod_ctx``: fixed shapes (B, N, A, K), fixed config, one device."""

    def __init__(self, B: int, N: int, A: int, K: int, config: BayesODConfig = None, device: int = 0):
        self.lib = _cabi.load()
        self.config = config or BayesODConfig()
        self.B, self.N, self.A, self.K = int(B), int(N), int(A), int(K)
        self.Dmax = self.config.max_output_size
        self.device = device
        self._ctx = C.c_void_p()
        self._ccfg = self.config.to_c(self.B, self.N, self.A, self.K)
        rc = self.lib.bod_create(C.byref(self._ctx), int(device), C.byref(self._ccfg))
        if rc != _cabi.BOD_OK:
            raise BodError(rc, (self.lib.bod_last_error(None) or b"").decode())
        self._alloc_host()

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.bod_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._ring = None
            for base in getattr(self, "_pinned_bases", []):
                self.lib.bod_host_free(base)
            self._pinned_bases = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _cabi.BOD_OK:
            raise BodError(rc, (self.lib.bod_last_error(self._ctx) or b"").decode())

    def _alloc_host(self):
        B, D, K = self.B, self.Dmax, self.K
        self._h = dict(num_dets=np.zeros(B, np.int32), num_survivors=np.zeros(B, np.int32),
                       means=np.zeros((B, D, 4), np.float32), covs=np.zeros((B, D, 4, 4), np.float32),
                       cat_param=np.zeros((B, D, K), np.float32), cat_count=np.zeros((B, D, K), np.float32),
                       nms_indices=np.zeros((B, D), np.int32), centre_anchor_idx=np.zeros((B, D), np.int32),
                       centre_scores=np.zeros((B, D), np.float32))
        self._hres = BodHostResults(**{k: v.ctypes.data for k, v in self._h.items()})

    def _results(self) -> Results:
        return Results(**{k: v.copy() for k, v in self._h.items()})

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.bod_workspace_bytes(self._ctx))

    @property
    def cov_width(self) -> int:
        return {0: 0, 1: 16, 2: 10}[self.config.cov_layout]

    # -- device path -------------------------------------------------------
    def run(self, cls, box, cov=None, anchors=None, counts=None, stream=0):
        """Asynchronous launch on ``stream`` (a cudaStream_t handle, 0 = default).
        All arguments are device arrays: cls [B,N,A,K], box [B,N,A,4],
        cov [B,N,A,4,4] | [B,N,A,10] | None, anchors [A,4] | None, counts [B,A,K] | None."""
        B, N, A, K = self.B, self.N, self.A, self.K
        keep = []
        p_cls = device_ptr(cls, B * N * A * K, keep)
        p_box = device_ptr(box, B * N * A * 4, keep)
        p_cov = device_ptr(cov, B * N * A * self.cov_width, keep) if self.cov_width else None
        p_anc = device_ptr(anchors, A * 4, keep) if self.config.anchor_mode == _cabi.ANCHORS_TENSOR else None
        p_cnt = device_ptr(counts, B * A * K, keep) if counts is not None else None
        # keep the buffers alive while any run that reads them may be in flight (one per lane)
        self._keeps = getattr(self, "_keeps", [])
        self._keeps.append((cls, box, cov, anchors, counts, keep))
        del self._keeps[:-max(1, int(self.config.pipeline_depth))]
        self._check(self.lib.bod_run(self._ctx, p_cls, p_box, p_cov, p_anc, p_cnt, C.c_void_p(int(stream) or None)))

    def run_levels(self, cls, box, cov=None, anchors=None, counts=None, stream=0):
        """bod_run_levels: ``cls`` / ``box`` / ``cov`` are sequences with one device array per FPN level
        ([B,N,A_l,K], [B,N,A_l,4], [B,N,A_l,4,4] | [B,N,A_l,10]) -- the head outputs before the reference
        concatenates them (retinanet_model.py:89-112)."""
        B, N, K = self.B, self.N, self.K
        la = list(self.config.level_anchors)
        keep = []
        ptrs = lambda seq, width: (C.c_void_p * len(la))(*[device_ptr(t, B * N * a * width, keep) for t, a in zip(seq, la)])   # noqa: E731
        if len(cls) != len(la) or len(box) != len(la) or (self.cov_width and len(cov) != len(la)):
            raise ValueError(f"expected {len(la)} tensors per kind (one per level)")
        p_cls, p_box = ptrs(cls, K), ptrs(box, 4)
        p_cov = ptrs(cov, self.cov_width) if self.cov_width else None
        p_anc = device_ptr(anchors, self.A * 4, keep) if self.config.anchor_mode == _cabi.ANCHORS_TENSOR else None
        p_cnt = device_ptr(counts, B * self.A * K, keep) if counts is not None else None
        self._keeps = getattr(self, "_keeps", [])
        self._keeps.append((cls, box, cov, anchors, counts, keep))
        del self._keeps[:-max(1, int(self.config.pipeline_depth))]
        self._check(self.lib.bod_run_levels(self._ctx, p_cls, p_box, p_cov, p_anc, p_cnt, C.c_void_p(int(stream) or None)))

    def validate(self, cls, box, anchors, scaling=None, stream=0):
        """validation_utils.post_process_predictions for the batch (bod_validate_run): device arrays
        cls [B,A,K], box [B,A,4], anchors [A,4]; ``scaling`` = None | (mode, shift[4], (norm_h, norm_w),
        (scale_h, scale_w)).  Fetch with ``fetch()``: classes in ``cat_param``, corners (vuvu) in ``means``."""
        B, A, K = self.B, self.A, self.K
        keep = []
        p_cls = device_ptr(cls, B * A * K, keep)
        p_box = device_ptr(box, B * A * 4, keep)
        p_anc = device_ptr(anchors, A * 4, keep)
        sc = None
        if scaling is not None:
            mode, shift, norm_hw, scale_hw = scaling
            sc = C.byref(_cabi.BodValScaling(int(mode), (C.c_float * 4)(*[float(x) for x in shift]), float(norm_hw[0]),
                                             float(norm_hw[1]), float(scale_hw[0]), float(scale_hw[1])))
        self._keep = (cls, box, anchors, keep)
        self._check(self.lib.bod_validate_run(self._ctx, p_cls, p_box, p_anc, sc, C.c_void_p(int(stream) or None)))

    def set_sampler_stream(self, seed: int, image_id_base: int):
        """Philox stream of the next runs: image b draws from (seed, image_id_base + b, anchor)."""
        self._check(self.lib.bod_set_sampler_stream(self._ctx, int(seed), int(image_id_base) & 0xFFFFFFFF))

    def set_input_hold(self, enabled: bool = True):
        """Promise (pipelined contexts) that every input of a run stays untouched until its results are complete; runs
        then do not wait for each other through the caller's stream (bod_set_input_hold)."""
        self._check(self.lib.bod_set_input_hold(self._ctx, 1 if enabled else 0))

    def set_image_scale(self, scale_v: float, scale_u: float):
        """KITTI rescale factors of the next runs (inference_utils.py:147-167)."""
        self._check(self.lib.bod_set_image_scale(self._ctx, float(scale_v), float(scale_u)))

    def wait_results(self, stream=0):
        """Make ``stream`` wait for the results of the last run (pipeline_depth = 2 only; bod_wait_results)."""
        self._check(self.lib.bod_wait_results(self._ctx, C.c_void_p(int(stream) or None)))

    def fetch(self) -> Results:
        self._check(self.lib.bod_fetch(self._ctx, C.byref(self._hres)))
        return self._results()

    # -- streaming retrieval: every run's results, not only the last one's -------------------
    @property
    def last_ticket(self) -> int:
        """Ticket of the run issued last (1, 2, 3, ...): bod_last_ticket."""
        return int(self.lib.bod_last_ticket(self._ctx))

    def _pinned_block(self):
        """One page-locked host copy of a lane's result block (bod_host_alloc), viewed as numpy arrays at the
        offsets of bod_result_block_layout."""
        B, D, K = self.B, self.Dmax, self.K
        offs, total = (C.c_int64 * 10)(), C.c_int64()
        self._check(self.lib.bod_result_block_layout(self._ctx, offs, C.byref(total)))
        shapes = [("num_dets", (B,), np.int32), ("num_survivors", (B,), np.int32), ("means", (B, D, 4), np.float32),
                  ("covs", (B, D, 4, 4), np.float32), ("cat_param", (B, D, K), np.float32), ("cat_count", (B, D, K), np.float32),
                  ("nms_indices", (B, D), np.int32), ("centre_anchor_idx", (B, D), np.int32), ("centre_scores", (B, D), np.float32)]
        base = self.lib.bod_host_alloc(int(total.value))
        if not base:
            raise MemoryError("bod_host_alloc failed")
        self._pinned_bases = getattr(self, "_pinned_bases", [])
        self._pinned_bases.append(base)
        arrays = {}
        for i, (k, sh, dt) in enumerate(shapes):
            n = int(np.prod(sh))
            buf = (C.c_byte * (n * 4)).from_address(base + int(offs[i]))
            arrays[k] = np.frombuffer(buf, dtype=dt, count=n).reshape(sh)
        return arrays, base

    def fetch_async(self, ticket: int = None) -> int:
        """Enqueue the device->host copy of run ``ticket`` (default: the last one) behind that run, into the
        pinned result ring slot of its lane (one cudaMemcpyAsync of the lane's result block:
        bod_fetch_block_async).  Nothing blocks; collect with ``collect(ticket)`` before ``pipeline_depth``
        further tickets have been fetched (the slot is then reused)."""
        ticket = int(self.lib.bod_last_ticket(self._ctx)) if ticket is None else int(ticket)
        ring = getattr(self, "_ring", None)
        if ring is None:
            ring = self._ring = [self._pinned_block() for _ in range(max(1, int(self.config.pipeline_depth)))]
        rc = self.lib.bod_fetch_block_async(self._ctx, ticket, ring[ticket % len(ring)][1])
        if rc != _cabi.BOD_OK:
            self._check(rc)
        return ticket

    def collect(self, ticket: int, copy=True):
        """Wait for run ``ticket`` and the copies ``fetch_async`` enqueued for it; its Results."""
        self._check(self.lib.bod_ticket_wait(self._ctx, int(ticket)))
        arrays, _ = self._ring[int(ticket) % len(self._ring)]
        return Results(**{k: v.copy() for k, v in arrays.items()}) if copy else arrays

    def device_results_at(self, ticket: int) -> dict:
        """Device addresses of run ``ticket``'s padded result blocks (bod_device_results_at)."""
        out = _cabi.BodDeviceResults()
        self._check(self.lib.bod_device_results_at(self._ctx, int(ticket), C.byref(out)))
        return {name: getattr(out, name) for name, _ in out._fields_}

    def fetch_into_pinned(self):
        """bod_fetch into the engine's own host blocks without copying them out (bench)."""
        self._check(self.lib.bod_fetch(self._ctx, C.byref(self._hres)))
        return self._h

    def synchronize(self):
        self._check(self.lib.bod_synchronize(self._ctx))

    # -- host path ---------------------------------------------------------
    @staticmethod
    def _host(a, n):
        a = np.ascontiguousarray(a, np.float32)
        if a.size != n:
            raise ValueError(f"array has {a.size} elements, expected {n}")
        return a

    def run_host(self, cls, box, cov=None, anchors=None, counts=None, copy=True) -> Results:
        """Synchronous call with numpy inputs (H2D staging inside the library)."""
        B, N, A, K = self.B, self.N, self.A, self.K
        a_cls = self._host(cls, B * N * A * K)
        a_box = self._host(box, B * N * A * 4)
        a_cov = self._host(cov, B * N * A * self.cov_width) if self.cov_width else None
        a_anc = self._host(anchors, A * 4) if self.config.anchor_mode == _cabi.ANCHORS_TENSOR else None
        a_cnt = self._host(counts, B * A * K) if counts is not None else None
        ptr = lambda x: x.ctypes.data if x is not None else None   # noqa: E731
        self._check(self.lib.bod_run_host(self._ctx, ptr(a_cls), ptr(a_box), ptr(a_cov), ptr(a_anc), ptr(a_cnt),
                                          C.byref(self._hres)))
        return self._results() if copy else self._h

    def run_host_ptrs(self, p_cls, p_box, p_cov, p_anc, p_cnt):
        """bod_run_host on raw host addresses (e.g. pinned torch tensors); results stay in the engine's blocks."""
        self._check(self.lib.bod_run_host(self._ctx, p_cls, p_box, p_cov, p_anc, p_cnt, C.byref(self._hres)))
        return self._h

    def host_traffic(self) -> dict:
        """Bytes the last run_host moved (copied H2D, gathered in place from pinned memory, copied D2H)."""
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self.lib.bod_last_host_traffic(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return dict(h2d_copied=a.value, h2d_gathered=b.value, d2h=c.value)

    def cluster_host(self, counts, means, covs, centres, affinity, affinity_threshold):
        """bayes_od_clustering for one image from host arrays (bod_cluster_host)."""
        counts = np.ascontiguousarray(counts, np.float32)
        S = counts.shape[0]
        means = np.ascontiguousarray(means, np.float32).reshape(S, 4)
        covs = np.ascontiguousarray(covs, np.float32).reshape(S, 16)
        centres = np.ascontiguousarray(centres, np.int32).reshape(-1)
        D = centres.shape[0]
        affinity = np.ascontiguousarray(affinity, np.float32).reshape(S, S)
        self._check(self.lib.bod_cluster_host(self._ctx, S, counts.ctypes.data, means.ctypes.data, covs.ctypes.data,
                                              D, centres.ctypes.data, affinity.ctypes.data,
                                              C.c_float(affinity_threshold), C.byref(self._hres)))
        h = self._h
        return (h["cat_param"][0, :D].copy(), h["means"][0, :D, :, None].copy(), h["covs"][0, :D].copy(),
                h["cat_count"][0, :D].copy())

    # -- parity intermediates ---------------------------------------------
    def survivors(self, b: int) -> dict:
        """The per-survivor outputs of bayes_od_inference for image b (inference_utils.py:217)."""
        cap, K = self.A if self.config.max_survivors <= 0 else min(self.A, self.config.max_survivors), self.K
        cap = (cap + 31) // 32 * 32
        out = dict(anchor_idx=np.zeros(cap, np.int32), counts=np.zeros((cap, K), np.float32),
                   means=np.zeros((cap, 4), np.float32), covs=np.zeros((cap, 4, 4), np.float32),
                   scores=np.zeros(cap, np.float32), corners=np.zeros((cap, 4), np.float32))
        hs = BodHostSurvivors(capacity=cap, count=0, **{k: v.ctypes.data for k, v in out.items()})
        self._check(self.lib.bod_fetch_survivors(self._ctx, b, C.byref(hs)))
        S = hs.count
        return {k: v[:S].copy() for k, v in out.items()}

    def members(self, b: int, S: int, D: int) -> np.ndarray:
        """[D, ceil(S/32)] uint32 membership bitmask of image b."""
        wpr = max((S + 31) // 32, 1)
        mask = np.zeros((max(D, 1), wpr), np.uint32)
        self._check(self.lib.bod_fetch_members(self._ctx, b, mask.ctypes.data, wpr))
        return mask[:D]

    def probs(self, b: int) -> np.ndarray:
        out = np.zeros((self.A, self.K), np.float32)
        self._check(self.lib.bod_fetch_probs(self._ctx, b, out.ctypes.data))
        return out

    def sampled_counts(self, b: int) -> np.ndarray:
        out = np.zeros((self.A, self.K), np.float32)
        self._check(self.lib.bod_fetch_sampled_counts(self._ctx, b, out.ctypes.data))
        return out

    def stage_ms(self) -> dict:
        ms = (C.c_float * 6)()
        self._check(self.lib.bod_last_stage_ms(self._ctx, ms))
        names = ["moments_filter", "scan", "posterior", "soft_nms", "fusion", "total"]
        return dict(zip(names, [float(x) for x in ms]))

    STAGES = ["moments_filter", "scan", "posterior", "soft_nms", "fusion", "total"]

    def set_stage_timing(self, enabled: bool):
        self._check(self.lib.bod_set_stage_timing(self._ctx, int(enabled)))

    def stage_ms_accum(self):
        """(dict of summed stage milliseconds, number of runs) since the previous call."""
        ms = (C.c_float * 6)()
        runs = C.c_int32(0)
        self._check(self.lib.bod_stage_ms_accum(self._ctx, ms, C.byref(runs)))
        return dict(zip(self.STAGES, [float(x) for x in ms])), int(runs.value)

    def moments_clock_accum(self):
        """(summed milliseconds, launches) of the moments kernel since the previous call, by the kernel's own launch
        clock (bod_moments_clock_accum)."""
        ms = C.c_double(0.0)
        runs = C.c_int32(0)
        self._check(self.lib.bod_moments_clock_accum(self._ctx, C.byref(ms), C.byref(runs)))
        return float(ms.value), int(runs.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.bod_last_launch_count(self._ctx))
