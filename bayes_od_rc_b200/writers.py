"""Batched result writers: the ``np.save`` x 4 at the end of every iteration of the
reference's ``run_inference.py`` loop (:241-244), for a whole batch of results at once,
in native code on a few host threads (``bod_write_results_npy``).  The files are
byte-identical to ``numpy.save`` output, so ``offline_eval/*/compute_{ap,pdq,uncertainty_error}.py``
load them unchanged."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi
from ._cabi import BodError, BodHostResults


def save_batch(results, sample_ids, mean_dir, cov_dir, cat_param_dir, cat_count_dir, nthreads: int = 8) -> None:
    """``results``: an ``engine.Results`` (padded blocks of B images) or the dict of arrays
    ``BayesODEngine.fetch_into_pinned`` returns; ``sample_ids``: B file stems
    (``dataset_handler.sample_ids[counter]`` in the reference)."""
    get = (lambda k: results[k]) if isinstance(results, dict) else (lambda k: getattr(results, k))
    arrs = {k: np.ascontiguousarray(get(k), np.int32 if k == "num_dets" else np.float32)
            for k in ("num_dets", "means", "covs", "cat_param", "cat_count")}
    B, D = arrs["means"].shape[0], arrs["means"].shape[1]
    K = arrs["cat_param"].shape[2]
    if len(sample_ids) != B:
        raise ValueError(f"{len(sample_ids)} sample ids for {B} images")
    for d in (mean_dir, cov_dir, cat_param_dir, cat_count_dir):
        os.makedirs(d, exist_ok=True)
    res = BodHostResults(num_dets=arrs["num_dets"].ctypes.data, means=arrs["means"].ctypes.data,
                         covs=arrs["covs"].ctypes.data, cat_param=arrs["cat_param"].ctypes.data,
                         cat_count=arrs["cat_count"].ctypes.data)
    ids = (C.c_char_p * B)(*[str(s).encode() for s in sample_ids])
    enc = lambda p: os.fspath(p).encode()          # noqa: E731
    rc = _cabi.load().bod_write_results_npy(C.byref(res), B, D, K, enc(mean_dir), enc(cov_dir), enc(cat_param_dir),
                                            enc(cat_count_dir), ids, int(nthreads))
    if rc != _cabi.BOD_OK:
        raise BodError(rc, "bod_write_results_npy failed (directory missing or not writable?)")
