"""ctypes declarations for include/bayesod.h (the C ABI of libbayesod.so).

The library is loaded from bayes_od_rc_b200/lib/ (built in-tree by
``bayes_od_rc_b200.build``).  There is no fallback: if the shared object is
missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbayesod.so")

BOD_OK = 0
BOD_ERR_STATE = -4
STATUS_NAMES = {0: "BOD_OK", -1: "BOD_ERR_INVALID", -2: "BOD_ERR_CUDA", -3: "BOD_ERR_NOMEM",
                -4: "BOD_ERR_STATE", -5: "BOD_ERR_OVERFLOW"}

COV_NONE, COV_FULL16, COV_PACKED10 = 0, 1, 2
PRIOR_NONE, DIRICHLET_NON_INFORMATIVE, GAUSSIAN_ISOTROPIC = 0, 1, 1
RANK_SCORE, RANK_JOINT_ENTROPY = 0, 1
ANCHORS_TENSOR, ANCHORS_GENERATE = 0, 1


class BodConfig(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("A", C.c_int32), ("K", C.c_int32),
        ("cov_layout", C.c_int32), ("use_full_covar", C.c_int32),
        ("dirichlet_prior", C.c_int32), ("gaussian_prior", C.c_int32),
        ("isotropic_variance", C.c_float), ("ranking_method", C.c_int32),
        ("max_output_size", C.c_int32), ("iou_threshold", C.c_float), ("soft_nms_sigma", C.c_float),
        ("scale_v", C.c_float), ("scale_u", C.c_float), ("cov_calibration", C.c_float),
        ("num_draws", C.c_int32), ("seed", C.c_uint64), ("image_id_base", C.c_uint32),
        ("score_threshold", C.c_float), ("pre_nms_top_k", C.c_int32),
        ("anchor_mode", C.c_int32), ("im_h", C.c_int32), ("im_w", C.c_int32),
        ("max_survivors", C.c_int32), ("emit_probs", C.c_int32), ("pipeline_depth", C.c_int32),
        ("n_levels", C.c_int32), ("level_anchors", C.c_int32 * 8),
    ]


class BodValScaling(C.Structure):
    _fields_ = [("mode", C.c_int32), ("shift", C.c_float * 4), ("norm_h", C.c_float), ("norm_w", C.c_float),
                ("scale_h", C.c_float), ("scale_w", C.c_float)]


VAL_SCALE_NONE, VAL_SCALE_KITTI, VAL_SCALE_COCO = 0, 1, 2


class BodHostResults(C.Structure):
    _fields_ = [
        ("num_dets", C.c_void_p), ("num_survivors", C.c_void_p), ("means", C.c_void_p), ("covs", C.c_void_p),
        ("cat_param", C.c_void_p), ("cat_count", C.c_void_p), ("nms_indices", C.c_void_p),
        ("centre_anchor_idx", C.c_void_p), ("centre_scores", C.c_void_p),
    ]


class BodDeviceResults(C.Structure):
    _fields_ = BodHostResults._fields_


class BodHostSurvivors(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32), ("count", C.c_int32), ("anchor_idx", C.c_void_p), ("counts", C.c_void_p),
        ("means", C.c_void_p), ("covs", C.c_void_p), ("scores", C.c_void_p), ("corners", C.c_void_p),
    ]


# every symbol include/bayesod.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bod_abi_version": (C.c_int, []),
    "bod_status_string": (C.c_char_p, [C.c_int]),
    "bod_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(BodConfig)]),
    "bod_destroy": (None, [C.c_void_p]),
    "bod_last_error": (C.c_char_p, [C.c_void_p]),
    "bod_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "bod_run": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p]),
    "bod_run_levels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bod_wait_results": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bod_set_sampler_stream": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32]),
    "bod_set_image_scale": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "bod_validate_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(BodValScaling), C.c_void_p]),
    "bod_run_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.POINTER(BodHostResults)]),
    "bod_last_host_traffic": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bod_cluster_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_float, C.POINTER(BodHostResults)]),
    "bod_fetch": (C.c_int, [C.c_void_p, C.POINTER(BodHostResults)]),
    "bod_device_results_of": (C.c_int, [C.c_void_p, C.POINTER(BodDeviceResults)]),
    "bod_last_ticket": (C.c_int64, [C.c_void_p]),
    "bod_result_block_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bod_fetch_block_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "bod_fetch_async": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(BodHostResults)]),
    "bod_ticket_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "bod_device_results_at": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(BodDeviceResults)]),
    "bod_host_alloc": (C.c_void_p, [C.c_size_t]),
    "bod_host_free": (None, [C.c_void_p]),
    "bod_fetch_survivors": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(BodHostSurvivors)]),
    "bod_fetch_members": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "bod_fetch_probs": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "bod_fetch_sampled_counts": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "bod_synchronize": (C.c_int, [C.c_void_p]),
    "bod_set_input_hold": (C.c_int, [C.c_void_p, C.c_int]),
    "bod_last_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "bod_moments_clock_accum": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "bod_set_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "bod_stage_ms_accum": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "bod_last_launch_count": (C.c_int, [C.c_void_p]),
    "bod_write_results_npy": (C.c_int, [C.POINTER(BodHostResults), C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p,
                                        C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int32]),
    "bod_write_npy": (C.c_int, [C.c_char_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32]),
    "bod_bdd_json_open": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.POINTER(C.c_char_p), C.c_int32]),
    "bod_bdd_json_append": (C.c_int, [C.c_void_p, C.POINTER(BodHostResults), C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_char_p)]),
    "bod_bdd_json_close": (C.c_int, [C.c_void_p]),
    "bod_write_results_kitti_txt": (C.c_int, [C.POINTER(BodHostResults), C.c_int32, C.c_int32, C.c_int32, C.c_char_p,
                                              C.POINTER(C.c_char_p), C.c_int32]),
    "bod_format_float": (C.c_int, [C.c_double, C.c_int32, C.c_char_p, C.c_int32]),
    "bod_pdq_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32]),
    "bod_pdq_destroy": (None, [C.c_void_p]),
    "bod_pdq_last_error": (C.c_char_p, [C.c_void_p]),
    "bod_pdq_heatmaps": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "bod_pdq_losses": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 8),
    "bod_pdq_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bod_pdq_bvn_cdf": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bod_entropies": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bod_mu_error": (C.c_int, [C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                               C.c_void_p]),
    "bod_uncertainty_last_error": (C.c_char_p, []),
    "bod_generate_anchors": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libbayesod.so and bind every declared symbol. Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bayes_od_rc_b200.build` "
                "(there is no CPU fallback for the BayesOD path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class BodError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
