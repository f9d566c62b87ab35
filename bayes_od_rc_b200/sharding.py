"""Multi-GPU layout of the path: images are independent units, so a batch is
split into contiguous shards, one per GPU / rank, with NO collective on the hot
path.  RNG counters and synthetic data are keyed by the
GLOBAL image id, so the bytes produced for image i do not depend on the shard
layout.  The only optional exchange is an all-gather of the fixed-size padded
detection blocks (~17 KB per image) for consumers that want every detection on
every rank; it runs after the last kernel of a shard and is bandwidth-trivial,
so it is a plain ``torch.distributed.all_gather`` (NCCL over NVLink on GPUs,
gloo in the CPU tests), not a fused kernel."""
from __future__ import annotations

RESULT_KEYS = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices",
               "centre_anchor_idx", "centre_scores")


def image_shard(total_images: int, world: int, rank: int):
    """Contiguous shard [first, first+count) of `total_images` for `rank`:
    ceil(total/world) images per rank, the tail ranks may get fewer (or none)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-total_images // world)
    first = min(rank * per, total_images)
    return first, min(per, total_images - first)


def allgather_detections(blocks: dict, world: int, group=None) -> dict:
    """All-gather the padded result blocks ([B_local, ...] tensors) of every rank and
    concatenate them in rank order.  Every rank must hold the same B_local
    (pad the last shard); rows of padded images have num_dets = 0."""
    import torch
    import torch.distributed as dist
    out = {}
    for k in RESULT_KEYS:
        if k not in blocks:
            continue
        t = blocks[k].contiguous()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        out[k] = torch.cat(parts, dim=0)
    return out


class _CudaBlock:
    """Zero-copy view of a device result block for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_result_tensors(engine) -> dict:
    """The engine's device-resident padded result blocks as torch tensors (no copy)."""
    import ctypes as C
    import torch
    from ._cabi import BodDeviceResults
    dr = BodDeviceResults()
    engine._check(engine.lib.bod_device_results_of(engine._ctx, C.byref(dr)))
    B, D, K = engine.B, engine.Dmax, engine.K
    shapes = dict(num_dets=((B,), "<i4"), num_survivors=((B,), "<i4"), means=((B, D, 4), "<f4"),
                  covs=((B, D, 4, 4), "<f4"), cat_param=((B, D, K), "<f4"), cat_count=((B, D, K), "<f4"),
                  nms_indices=((B, D), "<i4"), centre_anchor_idx=((B, D), "<i4"), centre_scores=((B, D), "<f4"))
    dev = torch.device("cuda", engine.device)
    return {k: torch.as_tensor(_CudaBlock(getattr(dr, k), shp, ts), device=dev) for k, (shp, ts) in shapes.items()}
