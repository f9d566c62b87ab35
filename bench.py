#!/usr/bin/env python
"""bench.py — BayesOD post-head throughput (head outputs -> fused detections).

    python bench.py --gpus 1 --steps 50 --warmup 5                 # our CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the CPU implementation of the path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W     # N GPUs, one rank each

A "step" is one pass of the hot path (inference_utils.py:25-217 + :285-364 of the
reference, minus the model call) over one batch of B synthetic BDD-shape images
per GPU: K1 moments/sampler/filter -> scan -> K2 posterior -> K3 soft-NMS ->
K4 fusion.  Images are sharded over GPUs by batch (weak scaling: B per GPU), no
collective on the hot path.  Rank 0 prints ONE JSON line.

  value     images/s over all GPUs, inputs resident in HBM, in-kernel Philox sampler
  e2e       images/s through the C-ABI host entry (bod_run_host): inputs in pinned
            host memory, H2D + compute + D2H of the result blocks inside the timed region
  roofline  K1 (the only kernel that touches every anchor): algorithmic bytes of one
            launch / its mean CUDA-event duration over the timed region, vs the
            measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port of the reference path on the host cores (N=1 only)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (im_h, im_w, N, K, B per GPU, use_full_covar, config_id)
    # BASELINE.json target: "batch of 32 BDD-shape images at N=10", "10 classes" -> K = 11, full covariance
    "bdd_covar_b32_k11": dict(im_h=720, im_w=1280, N=10, K=11, B=32, use_full_covar=True, config_id=3),
    "bdd_covar_b32_k8": dict(im_h=720, im_w=1280, N=10, K=8, B=32, use_full_covar=True, config_id=3),
    "bdd_kendall_b8_k8": dict(im_h=720, im_w=1280, N=10, K=8, B=8, use_full_covar=False, config_id=2),
    "kitti_covar_b64_n20_k4": dict(im_h=512, im_w=1696, N=20, K=4, B=64, use_full_covar=True, config_id=4,
                                   scale_v=375 / 512, scale_u=1242 / 1696),
    # BASELINE.json config 5 (clustering stress): threshold 0.01, top-k 10k per image, N=40; 128 images over 8 GPUs
    "stress_b16_n40_k11": dict(im_h=720, im_w=1280, N=40, K=11, B=16, use_full_covar=True, config_id=5,
                               spec=dict(g_min=80, g_max=120, fg_iou=0.2, fg_logit=1.0, bg_logit_for_fg=0.0, stray_frac=0.02),
                               score_threshold=0.01, pre_nms_top_k=10000),
    # BASELINE.json config 4, second shape: raw KITTI frames (375x1242, A = 88 398), no resize => no rescale
    "kitti_raw_b64_n20_k4": dict(im_h=375, im_w=1242, N=20, K=4, B=64, use_full_covar=True, config_id=6),
    # BASELINE.json config 1 shape on the GPU: one image per call, as run_inference.py drives the path (:68, :137-149)
    "bdd_covar_b1_k8": dict(im_h=720, im_w=1280, N=10, K=8, B=1, use_full_covar=True, config_id=1),
    "tiny": dict(im_h=192, im_w=320, N=10, K=8, B=4, use_full_covar=True, config_id=9),
}
DEFAULT_WORKLOAD = "bdd_covar_b32_k11"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the bench runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples = []            # (t, sm_mhz, reasons_bitmask, power_w)
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = float("nan")
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
                 0x10: "sync_boost"}
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "note": "no NVML samples"}
        sel = [s for s in self.samples if t0 <= s[0] <= t1]
        note = "sampled during the timed region"
        if len(sel) < 3:
            sel = self.samples
            note = "timed region shorter than the sampling period: samples span warm-up + timed + e2e"
        clocks = sorted(s[1] for s in sel)
        bits = 0
        for s in sel:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "power_w_max": max(s[3] for s in sel), "samples": len(sel), "note": note}


def default_lanes(B):
    """Lanes of the pipelined context for B images per step: the tails of a short run outlast several moments
    kernels (measured, round 2: B = 4: 8 lanes 49.7 k images/s, 12 lanes 52.3 k; B = 16: 4 lanes 61.5 k, 8 lanes 64.5 k;
    B = 32: 4, 6 and 8 lanes give the same)."""
    return 12 if 2 <= B <= 4 else (8 if B <= 16 else 4)


def workload_string(name, wl, A):
    """config.workload: the same string from both arms (ours / --impl reference)."""
    extra = ""
    if wl.get("pre_nms_top_k") or wl.get("score_threshold") is not None and wl.get("score_threshold", float("-inf")) > float("-inf"):
        extra = f" score_threshold={wl.get('score_threshold')} pre_nms_top_k={wl.get('pre_nms_top_k', 0)}"
    return (f"{name}: {wl['im_h']}x{wl['im_w']} A={A} N={wl['N']} K={wl['K']} B={wl['B']} images per step and GPU "
            f"use_full_covar={wl['use_full_covar']} cov=[N,A,4,4] philox-sampler seed=1234{extra}")


def bytes_min_per_image(N, A, K, S_mean, D_mean, cov_width, injected_counts=False, anchors_gathered=True):
    """SURVEY.md §8(d) / BASELINE.md §3 contract figure."""
    b = 4 * N * A * K + 4 * N * S_mean * (4 + cov_width) + (16 * S_mean if anchors_gathered else 0)
    if injected_counts:
        b += 4 * A * K
    b += 4 * D_mean * (4 + 16 + 2 * K) + 4
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override images per GPU")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (0 = auto)")
    ap.add_argument("--pipeline", type=int, default=0, choices=list(range(0, 17)),
                    help="bod_config.pipeline_depth (lanes): step i+1's K1/K2 overlap the soft-NMS/fusion of the steps "
                         "before it; 1 = one step at a time; 0 = auto: 4 lanes once a batch fills the GPU "
                         "(soft-NMS holds one SM per image), up to 8 for small batches")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-verify", action="store_true", help="skip the bit-for-bit check of timed results against the oracle")
    ap.add_argument("--no-input-hold", action="store_true", help="do not declare the inputs held (bod_set_input_hold)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (global batch split over the GPUs)")
    ap.add_argument("--no-stream-fetch", action="store_true", help="do not copy every step's result blocks to the host")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and world != args.gpus:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")

    if args.impl == "reference":
        return reference_arm(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    from bayes_od_rc_b200 import _cabi, synthetic
    from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    # keep this rank (and the pinned host buffers it will first-touch for the end-to-end leg) on the CPUs /
    # NUMA node its GPU hangs off: with 8 ranks the host side of the PCIe copies is what limits e2e
    numa_note = None
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa_note = f"rank pinned to the GPU's CPU set ({len(os.sched_getaffinity(0))} cpus)"
    except Exception as e:  # pragma: no cover
        numa_note = f"cpu affinity not set: {e!r}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = dict(WORKLOADS[args.workload])
    B = args.batch or wl["B"]
    wl["B"] = B
    N, K = wl["N"], wl["K"]
    if args.pipeline == 0:
        args.pipeline = default_lanes(B)
    spec = synthetic.SceneSpec(im_h=wl["im_h"], im_w=wl["im_w"], N=N, K=K, config_id=wl["config_id"], **wl.get("spec", {}))
    first_image = rank * B                      # global image ids: shard-invariant RNG + data

    # ---- synthetic head outputs, generated on the device, then resident in HBM ----
    t_gen = time.time()
    batch = synthetic.make_batch(spec, B, device=dev, with_counts=False, first_image_id=first_image)
    cls, box, cov, anchors = batch["cls"], batch["box"], batch["cov"], batch["anchors"]
    A = anchors.shape[0]
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    cfg = BayesODConfig(use_full_covar=wl["use_full_covar"], cov_layout=_cabi.COV_FULL16, seed=1234,
                        image_id_base=first_image, scale_v=wl.get("scale_v", 1.0), scale_u=wl.get("scale_u", 1.0),
                        max_survivors=min(A, 32768), pipeline_depth=args.pipeline,
                        score_threshold=wl.get("score_threshold", float("-inf")), pre_nms_top_k=wl.get("pre_nms_top_k", 0))
    eng = BayesODEngine(B, N, A, K, cfg, device=local_rank)
    if args.pipeline > 1 and not args.no_input_hold:
        eng.set_input_hold(True)                # the input tensors of a step are never overwritten here (see config.inputs)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream_fetch = not args.no_stream_fetch

    def step():
        # one pass of the path over the batch; every step's padded result blocks (576 KB at B = 32) are copied to
        # pinned host memory behind its own tail (bod_fetch_async), as a streaming consumer would take them
        eng.run(cls, box, cov, anchors, None, stream=stream.cuda_stream)
        if stream_fetch:
            eng.fetch_async()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident throughput ----
    # Set-up, before the warm-up steps: a pipelined context captures and instantiates the CUDA graphs of a lane on the
    # lane's second run, so every lane is taken through two runs first (config.graph_priming_steps; with the driver's
    # --warmup 5 three of the four captures would otherwise fall into a timed region of 20 steps).
    priming = 2 * args.pipeline if args.pipeline > 1 else 0
    for _ in range(priming):
        step()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.stage_ms_accum()                         # reset the stage accumulators
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    eng.wait_results(stream.cuda_stream)         # pipelined: the stream has only waited for the heads so far
    ev1.record(stream)
    barrier()
    t1 = time.perf_counter()
    elapsed_ms = ev0.elapsed_time(ev1)
    stage_sum, stage_runs = eng.stage_ms_accum()
    res = eng.collect(eng.last_ticket) if stream_fetch else eng.fetch()
    d2h_stream = int(sum(v.nbytes for v in eng._h.values())) if stream_fetch else 0
    S_mean = float(res.num_survivors.mean())
    D_mean = float(res.num_dets.mean())
    launches_per_step = eng.launch_count

    # ---- for reference: the same steps issued one after the other (no overlap between steps) ----
    serial = None
    if args.pipeline > 1:
        import dataclasses
        eng1 = BayesODEngine(B, N, A, K, dataclasses.replace(cfg, pipeline_depth=1), device=local_rank)
        n1 = max(3, min(args.steps, 20))
        for _ in range(3):
            eng1.run(cls, box, cov, anchors, None, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        eng1.stage_ms_accum()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(n1):
            eng1.run(cls, box, cov, anchors, None, stream=stream.cuda_stream)
        s1.record(stream)
        torch.cuda.synchronize()
        ssum, sruns = eng1.stage_ms_accum()
        serial = {"ms_per_step": round(s0.elapsed_time(s1) / n1, 4), "steps": n1,
                  "stage_ms": {k: round(v / max(sruns, 1), 4) for k, v in ssum.items()}}
        eng1.close()

    # ---- end to end: pinned host inputs -> bod_run_host -> host results ----
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.cpu().pin_memory()     # noqa: E731
        h_cls, h_box, h_cov, h_anc = pin(cls), pin(box), pin(cov), pin(anchors)
        n_e2e = args.e2e_steps or max(2, min(args.steps, 5))
        run_e2e = lambda: eng.run_host_ptrs(h_cls.data_ptr(), h_box.data_ptr(), h_cov.data_ptr(), h_anc.data_ptr(), None)   # noqa: E731
        run_e2e()                                 # warm-up (allocates the device staging buffers)
        barrier()
        te0 = time.perf_counter()
        for _ in range(n_e2e):
            out = run_e2e()
        torch.cuda.synchronize()
        te = time.perf_counter() - te0
        tr = eng.host_traffic()
        e2e = dict(seconds=te, steps=n_e2e, h2d=tr["h2d_copied"] + tr["h2d_gathered"], d2h=tr["d2h"],
                   h2d_copied=tr["h2d_copied"], h2d_gathered=tr["h2d_gathered"],
                   host_bytes=(cls.numel() + box.numel() + cov.numel() + anchors.numel()) * 4)
    # ---- strong scaling: the workload's batch as ONE global batch split over the GPUs (BASELINE.json configs 3 / 4:
    # "batch 32 ... image-sharded across 8xB200"): B/world images per GPU and step, streamed through the pipelined context
    strong = None
    if not args.no_strong and world > 1 and B % world == 0:
        Bg = B // world
        first_s = rank * Bg
        sb = synthetic.make_batch(spec, Bg, device=dev, with_counts=False, first_image_id=first_s)
        import dataclasses
        lanes_s = default_lanes(Bg)
        eng_s = BayesODEngine(Bg, N, A, K, dataclasses.replace(cfg, image_id_base=first_s, pipeline_depth=lanes_s), device=local_rank)
        if not args.no_input_hold:
            eng_s.set_input_hold(True)

        def step_s():
            eng_s.run(sb["cls"], sb["box"], sb["cov"], sb["anchors"], None, stream=stream.cuda_stream)
            if stream_fetch:
                eng_s.fetch_async()
        for _ in range(max(args.warmup, 3) + 2 * lanes_s):       # every lane through two runs (graph capture) + the warm-up steps
            step_s()
        barrier()
        z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        z0.record(stream)
        for _ in range(args.steps):
            step_s()
        eng_s.wait_results(stream.cuda_stream)
        z1.record(stream)
        barrier()
        strong = dict(ms=z0.elapsed_time(z1), images_per_gpu=Bg, lanes=lanes_s)
        eng_s.close()
        del sb

    # ---- the timed results against the oracle, bit for bit (outside every timed region) ----
    verification = None
    if not args.no_verify and rank == 0:
        verification = verify_against_oracle(cls, box, cov, anchors, wl, cfg, res, first_image, N, A, K, local_rank)

    # ---- what a bare pinned host->device copy reaches on this box (the e2e leg's ceiling) ----
    h2d_probe = None
    if e2e is not None:
        n_probe = min(cls.numel(), 256 << 20)
        src = h_cls.view(-1)[:n_probe]
        dst = torch.empty(n_probe, dtype=torch.float32, device=dev)
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        p1.record(); torch.cuda.synchronize()
        h2d_probe = 3 * n_probe * 4 / (p0.elapsed_time(p1) * 1e-3) / 1e9
        del dst
    sampler.stop()

    # ---- max over ranks ----
    if world > 1:
        tt = torch.tensor([elapsed_ms, e2e["seconds"] if e2e else 0.0, strong["ms"] if strong else 0.0], device=dev,
                          dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt[0]); e2e_seconds = float(tt[1])
        if strong:
            strong["ms"] = float(tt[2])
        ss = torch.tensor([S_mean, D_mean], device=dev, dtype=torch.float64)
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        S_mean, D_mean = float(ss[0]) / world, float(ss[1]) / world
    elif e2e:
        e2e_seconds = e2e["seconds"]

    if rank == 0:
        peak, peak_src = measured_peaks()
        images = world * B * args.steps
        value = images / (elapsed_ms * 1e-3)
        k1_ms = stage_sum["moments_filter"] / max(stage_runs, 1)
        k1_bytes = 4.0 * N * A * K * B            # algorithmic bytes of ONE K1 launch: the [B,N,A,K] logits, read once
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None
        traffic = None                            # dram bytes of one K1 launch from the committed ncu --set full capture
        tp = os.path.join(ROOT, "profiles", "k1_traffic_r2.json")
        if not os.path.exists(tp):
            tp = os.path.join(ROOT, "profiles", "k1_traffic_r1.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("workload") == args.workload and B == wl["B"]:
                traffic = int(tj["dram_bytes_per_launch"])
        cw = 16
        bmin = bytes_min_per_image(N, A, K, S_mean, D_mean, cw)
        path_frac = (bmin * B / (elapsed_ms / args.steps * 1e-3) / 1e9) / peak
        line = {
            "metric": "BayesOD images/s (head out -> fused dets)", "value": round(value, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, wl, A),
                       "images_per_gpu": B, "global_batch": B * world, "parallelism": f"image-shard x{world}, no collective",
                       "l2": "inputs (%.2f GB per step) larger than L2" % ((cls.numel() + box.numel() + cov.numel()) * 4 / 1e9),
                       "pipeline_depth": args.pipeline,
                       "graph_priming_steps": priming,
                       "inputs": ("resident in HBM, never overwritten: held until each run's results are complete "
                                  "(bod_set_input_hold), so consecutive steps do not depend on each other through the "
                                  "caller's stream") if (args.pipeline > 1 and not args.no_input_hold) else
                                 "resident in HBM; the caller's stream waits for each run's logits to be consumed",
                       "results": ("every step's padded result blocks copied to pinned host memory behind its tail "
                                   f"(bod_fetch_async, {d2h_stream} bytes per step)") if stream_fetch else "left on the device",
                       "mean_survivors": round(S_mean, 1), "mean_dets": round(D_mean, 1),
                       "bytes_min_per_image": int(bmin), "path_roofline_frac": round(path_frac, 4)},
            "gpu_launches": launches_per_step * args.steps,
            "stage_ms": {k: round(v / max(stage_runs, 1), 4) for k, v in stage_sum.items()},
            "roofline": {"bound": "hbm", "kernel": "k1_moments_pipe_kernel", "achieved": round(achieved, 1) if achieved else None,
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(k1_bytes), "launch_ms": round(k1_ms, 4),
                         "launches_timed": int(stage_runs),
                         "timer": ("the kernel's own launch clock (%globaltimer at the start of its first CTA and at the end of "
                                   "its last one), every launch of the timed region that the per-lane 64-slot clocks still hold; "
                                   "CUDA events around a kernel of a pipelined context would put ~10 us between two launches. "
                                   "tests/test_gpu_parity.py::test_moments_launch_clock_agrees_with_cuda_events checks it against "
                                   "CUDA events on a serial context") if args.pipeline > 1 else "CUDA events on the launching stream"},
            "clocks": sampler.summary(t0, t1),
        }
        if verification is not None:
            line["verified"] = bool(verification["ok"])
            line["verification"] = verification
        # weak scaling is the line's `value`; the strong block is the same workload's batch as one global batch
        if strong:
            sv = B * args.steps / (strong["ms"] * 1e-3)
            line["strong"] = {"global_batch": B, "images_per_gpu": strong["images_per_gpu"], "n_gpus": world,
                              "value": round(sv, 1), "unit": "images/s", "ms_per_step": round(strong["ms"] / args.steps, 4),
                              "pipeline_depth": strong["lanes"],
                              "one_gpu_reference": round(value / world, 1),
                              "speedup_vs_one_gpu": round(sv / (value / world), 3),
                              "note": "one_gpu_reference = this run's per-GPU throughput at the full batch per GPU (the weak-"
                                      "scaling `value` / n_gpus), i.e. what one GPU does with the whole global batch"}
        elif world == 1 and not args.no_strong:
            line["strong"] = {"global_batch": B, "images_per_gpu": B, "n_gpus": 1, "value": round(value, 1), "unit": "images/s",
                              "ms_per_step": round(elapsed_ms / args.steps, 4), "pipeline_depth": args.pipeline,
                              "one_gpu_reference": round(value, 1), "speedup_vs_one_gpu": 1.0,
                              "note": "one GPU: the strong- and weak-scaling runs are the same run"}
        if serial:
            line["serial"] = serial           # pipeline_depth = 1: whole steps back to back, stage times undisturbed
            k1_alone = serial["stage_ms"]["moments_filter"]
            if k1_alone > 0:                  # the same kernel with nothing running beside it (one step at a time)
                line["roofline"]["alone"] = {"launch_ms": k1_alone, "achieved": round(k1_bytes / (k1_alone * 1e-3) / 1e9, 1),
                                             "frac": round(k1_bytes / (k1_alone * 1e-3) / 1e9 / peak, 4),
                                             "note": "pipeline_depth=1; in the timed (pipelined) region the posterior, soft-NMS and "
                                                     "fusion kernels of earlier steps share the SMs with this launch"}
        if e2e:
            e2e_value = world * B * e2e["steps"] / e2e_seconds
            line["e2e"] = {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": int(e2e["h2d"]),
                           "d2h_bytes_per_step": int(e2e["d2h"]), "steps": e2e["steps"],
                           "h2d_copied": int(e2e["h2d_copied"]), "h2d_gathered_in_place": int(e2e["h2d_gathered"]),
                           "host_input_bytes": int(e2e["host_bytes"]), "host_placement": numa_note,
                           "h2d_gbs": round(e2e["h2d_copied"] * e2e["steps"] / e2e["seconds"] / 1e9, 1),
                           "h2d_probe_gbs": round(h2d_probe, 1) if h2d_probe else None,
                           "pcie_frac": round(e2e["h2d_copied"] * e2e["steps"] / e2e["seconds"] / 1e9 / h2d_probe, 3) if h2d_probe else None,
                           "api": "bod_run_host: pinned host buffers -> padded host result blocks; cls is copied in "
                                  "image chunks overlapped with compute, box/cov rows of the survivors are "
                                  "gathered in place from pinned memory"}
        # ---- CPU baseline: the oracle port on the host cores, bounded sample ----
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cls, box, cov, anchors, wl, args.cpu_sample, first_image)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def verify_against_oracle(cls, box, cov, anchors, wl, cfg, res, first_image, N, A, K, device, n=2):
    """Images 0..n-1 of the timed batch against the oracle, bit for bit (sampler included): a second context
    that keeps the mean probabilities and the sampled counts re-runs those images; its Philox counts must equal
    the oracle's restatement on the kernel's own probabilities (the one tolerance-checked quantity of the path),
    the oracle then runs the rest of the path on those counts, and BOTH the re-run and the timed run's last
    results must equal it exactly."""
    import dataclasses
    import numpy as np
    import oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    n = min(n, cls.shape[0])
    out = {"ok": False, "images": n, "checked": ["num_dets", "nms_indices", "centre_scores", "means", "covs", "cat_param", "cat_count"]}
    try:
        engv = BayesODEngine(n, N, A, K, dataclasses.replace(cfg, pipeline_depth=1, emit_probs=True, image_id_base=first_image),
                             device=device)
        engv.run(cls[:n].contiguous(), box[:n].contiguous(), cov[:n].contiguous(), anchors, None)
        rv = engv.fetch()
        oc = _oracle_cfg(wl, first_image)
        a = anchors.cpu().numpy()
        bad = []
        for b in range(n):
            counts = engv.sampled_counts(b)
            ref_counts = oracle.philox_counts(engv.probs(b), 30, 1234, first_image + b)
            if not np.array_equal(counts, ref_counts):
                bad.append(f"image {b}: sampled counts")
            r = oracle.run_image(oc, cls[b].cpu().numpy(), box[b].cpu().numpy(), cov[b].cpu().numpy(), a, counts,
                                 image_id=b, with_probs=False)
            D = len(r.nms_indices)
            for name, got in (("re-run", rv), ("timed run", res)):
                pairs = [("num_dets", int(got.num_dets[b]), D), ("nms_indices", got.nms_indices[b, :D], r.nms_indices),
                         ("centre_scores", got.centre_scores[b, :D], r.nms_scores), ("means", got.means[b, :D], r.final_means),
                         ("covs", got.covs[b, :D], r.final_covs), ("cat_param", got.cat_param[b, :D], r.final_scores),
                         ("cat_count", got.cat_count[b, :D], r.final_counts)]
                for key, x, y in pairs:
                    x = np.asarray(x); y = np.asarray(y)
                    if x.shape != y.shape or x.tobytes() != y.tobytes():
                        bad.append(f"image {b}, {name}: {key}")
            out.setdefault("dets", []).append(D)
            out.setdefault("survivors", []).append(int(len(r.keep)))
        engv.close()
        out["ok"] = not bad
        out["mismatches"] = bad
        out["how"] = ("oracle/bayesod_oracle.c on the same inputs; Philox counts checked against the restatement on the kernel's "
                      "own mean probabilities; every fused output compared byte for byte")
    except Exception as e:  # pragma: no cover
        out["error"] = repr(e)
    return out


def _oracle_cfg(wl, image_id_base):
    import oracle
    return oracle.OracleConfig(use_full_covar=wl["use_full_covar"], cov_layout=1, seed=1234, image_id_base=image_id_base,
                               scale_v=wl.get("scale_v", 1.0), scale_u=wl.get("scale_u", 1.0),
                               score_threshold=wl.get("score_threshold", float("-inf")),
                               pre_nms_top_k=wl.get("pre_nms_top_k", 0))


def cpu_baseline(cls, box, cov, anchors, wl, sample, first_image):
    """The oracle (CPU port of the reference path, Philox sampler included) on a
    bounded sample of the same images, all host cores (one image per thread)."""
    import oracle
    cores = os.cpu_count() or 1
    B = cls.shape[0]
    n = sample or min(B, max(4, min(cores, 32)))
    n = min(n, B)
    threads = min(cores, n)
    c = cls[:n].cpu().numpy(); b = box[:n].cpu().numpy(); v = cov[:n].cpu().numpy(); a = anchors.cpu().numpy()
    oc = _oracle_cfg(wl, first_image)
    oracle.run_batch(oc, c[:1], b[:1], v[:1], a, None, nthreads=1)     # warm-up (page in, build)
    # bounded: passes over the sample until ~1.5 s of wall clock (about 20 CPU-seconds on 16 threads), 2..20 passes
    passes, t0 = 0, time.perf_counter()
    while passes < 2 or (passes < 20 and time.perf_counter() - t0 < 1.5):
        oracle.run_batch(oc, c, b, v, a, None, nthreads=threads)
        passes += 1
    dt = time.perf_counter() - t0
    return {"value": round(n * passes / dt, 3), "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{n} images of the same workload x {passes} passes in {dt:.2f} s, oracle/bayesod_oracle.c (C "
                      f"restatement of inference_utils.py:25-217,285-364), {threads} host threads (host has {cores})"}


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The
    reference is Python/TensorFlow and TensorFlow is not installable offline, so
    this arm times the oracle port (oracle/bayesod_oracle.c) on the host cores."""
    if rank != 0:
        return
    import numpy as np
    import torch
    import oracle
    from bayes_od_rc_b200 import synthetic
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    N, K = wl["N"], wl["K"]
    cores = os.cpu_count() or 1
    n = wl["B"]                                   # one step = the workload's batch, as in the other arm
    threads = min(cores, n)
    spec = synthetic.SceneSpec(im_h=wl["im_h"], im_w=wl["im_w"], N=N, K=K, config_id=wl["config_id"], **wl.get("spec", {}))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    # the sample held in host memory: up to 16 distinct images, repeated to fill a step (the CPU cost per image is what
    # is measured; 32 BDD-shape images are 6.9 GB of host memory, 64 KITTI ones 20 GB)
    n_distinct = min(n, 16)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, n_distinct, device=dev, with_counts=False))
    oc = _oracle_cfg(wl, 0)
    A = batch["anchors"].shape[0]
    reps = -(-n // n_distinct)

    def run():
        done = 0
        for _ in range(reps):
            m = min(n_distinct, n - done)
            oracle.run_batch(oc, batch["cls"][:m], batch["box"][:m], batch["cov"][:m], batch["anchors"], None,
                             nthreads=min(threads, m))
            done += m
    warm = min(args.warmup, 1)
    for _ in range(warm):
        run()
    # exactly --steps steps unless that would take more than ~2 minutes (then as many as fit; the line says how many)
    steps, t0 = 0, time.perf_counter()
    while steps < max(1, args.steps) and (steps < 2 or time.perf_counter() - t0 < 120.0):
        run()
        steps += 1
    dt = time.perf_counter() - t0
    value = n * steps / dt
    line = {"impl": "reference", "metric": "BayesOD images/s (head out -> fused dets)", "value": round(value, 3),
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, wl, A), "images_per_step": n,
                       "distinct_images": n_distinct},
            "cpu_baseline": {"value": round(value, 3), "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{n} images per step x {steps} steps; the reference path is Python/TF (not "
                                       f"installable offline), timed: its C oracle port on {threads} host threads"},
            "e2e": {"value": round(value, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
