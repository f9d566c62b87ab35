"""Batched result writers: the ``np.save`` x 4 at the end of every iteration of the
reference's ``run_inference.py`` loop (:241-244), for a whole batch of results at once,
in native code on a few host threads (``bod_write_results_npy``).  The files are
byte-identical to ``numpy.save`` output, so ``offline_eval/*/compute_{ap,pdq,uncertainty_error}.py``
load them unchanged.

``BddJsonWriter`` and ``save_kitti_txt_batch`` do the same for the detection files of
``run_inference.py:176-212, 258-260`` (``predictions.json`` for bdd / coco / pascal,
one ``<id>.txt`` per image for kitti), also byte-identical to the reference's output."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi
from ._cabi import BodError, BodHostResults


def _host_results(results, keys):
    get = (lambda k: results[k]) if isinstance(results, dict) else (lambda k: getattr(results, k))
    arrs = {k: np.ascontiguousarray(get(k), np.int32 if k == "num_dets" else np.float32) for k in keys}
    res = BodHostResults(**{k: a.ctypes.data for k, a in arrs.items()})
    return arrs, res


def _ids(sample_ids):
    return (C.c_char_p * len(sample_ids))(*[str(s).encode() for s in sample_ids])


class BddJsonWriter:
    """``predictions.json`` of the reference's inference loop: ``append`` replaces
    ``final_results_list.extend(predictions_to_bdd_format(...))`` (``run_inference.py:206-212``,
    ``validation_utils.py:183-213``) for a result block of B images, ``close`` replaces the
    ``json.dump`` at ``:258-260``.  ``cat_param`` overrides the class block (the output of
    ``map_dataset_classes`` when training and test data sets differ)."""

    def __init__(self, path, category_list):
        self._h = C.c_void_p()
        cats = (C.c_char_p * max(len(category_list), 1))(*[str(c).encode() for c in category_list])
        rc = _cabi.load().bod_bdd_json_open(C.byref(self._h), os.fspath(path).encode(), cats, len(category_list))
        if rc != _cabi.BOD_OK:
            raise BodError(rc, f"cannot open {path}")

    def append(self, results, sample_ids, cat_param=None) -> None:
        if self._h is None:
            raise BodError(_cabi.BOD_ERR_STATE, "writer is closed")
        arrs, res = _host_results(results, ("num_dets", "means", "cat_param"))
        if cat_param is not None:
            arrs["cat_param"] = np.ascontiguousarray(cat_param, np.float32)
            res.cat_param = arrs["cat_param"].ctypes.data
        B, D = arrs["means"].shape[:2]
        if len(sample_ids) != B or arrs["cat_param"].shape[:2] != (B, D):
            raise ValueError("sample ids / class block do not match the result block")
        rc = _cabi.load().bod_bdd_json_append(self._h, C.byref(res), B, D, arrs["cat_param"].shape[2], _ids(sample_ids))
        if rc != _cabi.BOD_OK:
            raise BodError(rc, "bod_bdd_json_append failed")

    def close(self) -> None:
        if self._h is not None:
            h, self._h = self._h, None
            rc = _cabi.load().bod_bdd_json_close(h)
            if rc != _cabi.BOD_OK:
                raise BodError(rc, "bod_bdd_json_close failed")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def save_kitti_txt_batch(results, sample_ids, data_dir, cat_param=None, nthreads: int = 8) -> None:
    """``<data_dir>/<id>.txt`` for every image of a result block, as ``run_inference.py:176-201``
    writes ``predictions_to_kitti_format`` (``validation_utils.py:216-272``)."""
    arrs, res = _host_results(results, ("num_dets", "means", "cat_param"))
    if cat_param is not None:
        arrs["cat_param"] = np.ascontiguousarray(cat_param, np.float32)
        res.cat_param = arrs["cat_param"].ctypes.data
    B, D = arrs["means"].shape[:2]
    if len(sample_ids) != B or arrs["cat_param"].shape[:2] != (B, D):
        raise ValueError("sample ids / class block do not match the result block")
    os.makedirs(data_dir, exist_ok=True)
    rc = _cabi.load().bod_write_results_kitti_txt(C.byref(res), B, D, arrs["cat_param"].shape[2],
                                                  os.fspath(data_dir).encode(), _ids(sample_ids), int(nthreads))
    if rc != _cabi.BOD_OK:
        raise BodError(rc, "bod_write_results_kitti_txt failed (directory missing or not writable?)")


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
