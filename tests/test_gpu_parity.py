"""GPU parity tests proper: the CUDA path, called through the C ABI, against the
CPU oracle on the same inputs (bit-exact) and against the golden fixtures
(reference sources over the TF shim; tolerance rtol 1e-4 / atol 1e-5)."""
import numpy as np
import pytest

import oracle
from bayes_od_rc_b200 import anchors as anchors_mod
from bayes_od_rc_b200 import synthetic
from gpu_common import assert_bit_equal, compare_image_with_oracle, run_gpu_batch
from helpers import (adjudicated_close, check_categorical_merge, golden_cases, load_golden, oracle_config_of,
                     val_golden_cases, val_scaling_of, within_tol)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_cases())
def test_golden_fixture(name):
    g = load_golden(name)
    oc = oracle_config_of(g["meta"])
    cov = g["cov"] if g["meta"]["has_cov"] else None
    eng, res = run_gpu_batch(oc, g["cls"][None], g["box"][None], None if cov is None else cov[None], g["anchors"],
                             g["counts"][None])
    r = oracle.run_image(oc, g["cls"], g["box"], cov, g["anchors"], g["counts"])
    K = g["cls"].shape[-1]
    compare_image_with_oracle(eng, res, 0, r, K)
    # and against what the reference's own code produced
    S, D = len(g["cnt_post"]), len(g["nms_indices"])
    assert int(res.num_survivors[0]) == S and int(res.num_dets[0]) == D
    if S == 0:
        return
    sv = eng.survivors(0)
    assert np.array_equal(sv["counts"], g["cnt_post"])
    assert within_tol(sv["means"], g["mu_post"][:, :, 0]).all()
    assert np.array_equal(res.nms_indices[0, :D], g["nms_indices"])
    mem = oracle.mask_to_bool(eng.members(0, S, D), S)
    assert np.array_equal(mem, (g["iou_cols"] > oc.iou_threshold).T)
    assert within_tol(res.means[0, :D], g["final_means"][:, :, 0]).all()
    r64 = oracle.run_image(oc, g["cls"], g["box"], cov, g["anchors"], g["counts"], real="f64",
                           force=dict(nms_indices=r.nms_indices, mask=r.mask))
    ok, frac = adjudicated_close(res.covs[0, :D], g["final_covs"], r64.final_covs)
    assert ok.all() and frac > 0.99
    check_categorical_merge(g, g["nms_indices"], mem, r.extra["chosen"], res.cat_param[0, :D], res.cat_count[0, :D])


CONFIGS = {
    # name: (SceneSpec kwargs, OracleConfig kwargs, B)
    "bdd_covar_k8": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=3), dict(), 3),
    "bdd_kendall_k8": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=2),
                       dict(use_full_covar=False), 2),
    "bdd_covar_k11": (dict(im_h=192, im_w=320, N=10, K=11, g_min=6, g_max=10, box_hi=150., config_id=31), dict(), 2),
    "kitti_k4_n20": (dict(im_h=128, im_w=424, N=20, K=4, g_min=5, g_max=9, box_hi=120., config_id=4),
                     dict(scale_v=375 / 512, scale_u=1242 / 1696), 2),
    "packed_cov": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=33, packed_cov=True),
                   dict(cov_layout=2), 2),
    "n40_k11": (dict(im_h=96, im_w=160, N=40, K=11, g_min=4, g_max=6, box_hi=90., config_id=5), dict(), 2),
    "hard_nms": (dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=34),
                 dict(soft_nms_sigma=0.0), 2),
    "joint_entropy": (dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=35),
                      dict(ranking_method="joint_entropy"), 2),
    "no_gaussian_prior": (dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=36),
                          dict(gaussian_prior="None"), 1),
    # one big object: every survivor overlaps every other one, so soft-NMS candidates pile up dozens of
    # pending weights (exercises the replay of lazy commits and the pending-list overflow path)
    "dense_cluster": (dict(im_h=96, im_w=160, N=6, K=8, g_min=1, g_max=1, box_lo=70., box_hi=90., fg_iou=0.25, config_id=38),
                      dict(), 3),
    "dense_cluster_sigma": (dict(im_h=96, im_w=160, N=6, K=8, g_min=2, g_max=2, box_lo=60., box_hi=90., fg_iou=0.25, config_id=39),
                            dict(soft_nms_sigma=0.3, max_output_size=200), 2),
    "odd_A_k7": (dict(im_h=70, im_w=90, N=5, K=7, g_min=3, g_max=4, box_hi=60., config_id=37), dict(), 2),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_synthetic_batch_bit_exact(name):
    spec_kw, oc_kw, B = CONFIGS[name]
    spec = synthetic.SceneSpec(**spec_kw)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig(**oc_kw)
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"])
    for b in range(B):
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], batch["cov"][b], batch["anchors"], batch["counts"][b])
        assert len(r.keep) > 0
        compare_image_with_oracle(eng, res, b, r, spec.K)


@pytest.mark.parametrize("env", [dict(BOD_K3_PSM_MAX="0"), dict(BOD_K3_PSM_MAX="4"), dict(BOD_K3_THREADS="256"),
                                 dict(BOD_K3_THREADS="1024"), dict(BOD_K3_THREADS="256", BOD_K3_PSM_MAX="0"),
                                 dict(BOD_K3_SEGCAP="32"), dict(BOD_K3_SEGCAP="48", BOD_K3_PSM_MAX="0", BOD_K3_THREADS="1024")])
@pytest.mark.parametrize("name", ["bdd_covar_k8", "dense_cluster", "dense_cluster_sigma", "hard_nms"])
def test_softnms_variants(name, env, monkeypatch):
    """The soft-NMS kernel's other shapes on small inputs: pending weights that spill from shared memory to
    the global rows, the 256- / 1024-thread CTAs (other candidate-to-warp maps, other per-warp list lengths), and
    pair lists that overflow the warp's segment (a batch applied centre by centre / row by row)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    test_synthetic_batch_bit_exact(name)


PREFILTER = {
    # name: (SceneSpec kwargs, OracleConfig kwargs)
    "top_k": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=51), dict(pre_nms_top_k=200)),
    "top_k_k11": (dict(im_h=192, im_w=320, N=6, K=11, g_min=6, g_max=10, box_hi=150., fg_logit=1.0, bg_logit_for_fg=0.0,
                       stray_frac=0.02, config_id=52), dict(pre_nms_top_k=300)),
    "threshold": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., fg_logit=1.0, bg_logit_for_fg=0.0,
                       config_id=53), dict(score_threshold=0.6)),
    "both": (dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., fg_logit=1.0, bg_logit_for_fg=0.0,
                  config_id=54), dict(score_threshold=0.3, pre_nms_top_k=100)),
    "top_k_inactive": (dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=55),
                       dict(pre_nms_top_k=100000, score_threshold=0.01)),
    "drops_everything": (dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=56),
                         dict(score_threshold=2.0)),
}


@pytest.mark.parametrize("name", sorted(PREFILTER))
def test_prefilter_extension_bit_exact(name):
    """score_threshold / pre_nms_top_k (BASELINE config 5 knobs; semantics = the oracle's orc_prefilter):
    the kept anchors and everything downstream are bit-exact."""
    spec_kw, oc_kw = PREFILTER[name]
    spec = synthetic.SceneSpec(**spec_kw)
    B = 3
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig(**oc_kw)
    plain = oracle.OracleConfig()
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"])
    cut = 0
    for b in range(B):
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], batch["cov"][b], batch["anchors"], batch["counts"][b])
        r0 = oracle.category_filter(batch["counts"][b])
        cut += len(r0) - len(r.keep)
        compare_image_with_oracle(eng, res, b, r, spec.K)
    if name in ("top_k", "top_k_k11", "threshold", "both"):
        assert cut > 0, "the knob did not bite: the case tests nothing"
    if name == "drops_everything":
        assert res.num_survivors.sum() == 0 and res.num_dets.sum() == 0
    del plain


def test_prefilter_top_k_above_the_fast_softnms_capacity():
    """Stress shape (config 5): tens of thousands of survivors capped to 10 000 by pre_nms_top_k, which is
    more than the shared-memory soft-NMS kernel holds (falls through to the global-memory kernel)."""
    spec = synthetic.SceneSpec(N=4, K=11, g_min=80, g_max=120, fg_iou=0.2, fg_logit=1.0, bg_logit_for_fg=0.0,
                               stray_frac=0.02, config_id=5)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 1))
    oc = oracle.OracleConfig(pre_nms_top_k=10000, score_threshold=0.01)
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    r = oracle.run_image(oc, batch["cls"][0], batch["box"][0], batch["cov"][0], batch["anchors"], batch["counts"][0],
                         with_probs=False)
    assert len(oracle.category_filter(batch["counts"][0])) > 10000 and len(r.keep) == 10000
    compare_image_with_oracle(eng, res, 0, r, 11, check_probs=False)


@pytest.mark.parametrize("mode", ["big", "generic"])
@pytest.mark.parametrize("name", ["bdd_covar_k8", "dense_cluster_sigma", "hard_nms", "kitti_k4_n20"])
def test_softnms_large_survivor_variants(name, mode, monkeypatch):
    """The soft-NMS variants for survivor counts beyond the shared-memory pool (state in global memory) and
    beyond 16-bit list entries (the literal round-per-selection kernel), forced on small inputs."""
    monkeypatch.setenv("BOD_K3_MODE", mode)
    test_synthetic_batch_bit_exact(name)


def test_host_path_equals_device_path():
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=3)
    B = 9       # > 8 so that the host path splits the batch into chunks
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig()
    _, res_d = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    _, res_h = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"],
                             emit_probs=False, via_host=True)
    for k in ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx"):
        assert_bit_equal(getattr(res_h, k), getattr(res_d, k), k)


def test_pinned_host_path_gathers_in_place():
    """With pinned host buffers only `cls` is copied; box/cov rows of the survivors are read in place
    over PCIe.  Results must be identical and the reported traffic must show the saving."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=3)
    B = 9
    batch = synthetic.make_batch(spec, B)
    nb = synthetic.to_numpy(batch)
    oc = oracle.OracleConfig()
    _, res_d = run_gpu_batch(oc, nb["cls"], nb["box"], nb["cov"], nb["anchors"], nb["counts"], emit_probs=False)
    N, A, K = nb["cls"].shape[1:]
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc))
    pin = {k: batch[k].contiguous().pin_memory() for k in ("cls", "box", "cov", "anchors", "counts")}
    h = eng.run_host_ptrs(pin["cls"].data_ptr(), pin["box"].data_ptr(), pin["cov"].data_ptr(), pin["anchors"].data_ptr(),
                          pin["counts"].data_ptr())
    for k in ("num_dets", "num_survivors", "means", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx"):
        assert_bit_equal(h[k], getattr(res_d, k), k)
    assert_bit_equal(h["covs"], res_d.covs, "covs")
    tr = eng.host_traffic()
    full = (nb["cls"].size + nb["box"].size + nb["cov"].size + nb["counts"].size + nb["anchors"].size) * 4
    assert tr["h2d_gathered"] > 0
    assert tr["h2d_copied"] + tr["h2d_gathered"] < 0.6 * full          # box + cov tensors were not copied
    assert tr["h2d_copied"] == (nb["cls"].size + nb["counts"].size + nb["anchors"].size) * 4


def test_batch_composition_independence():
    """Image i's result does not depend on what else is in the batch."""
    spec = synthetic.SceneSpec(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=41)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 4))
    oc = oracle.OracleConfig()
    _, res4 = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    for b in (0, 3):
        _, res1 = run_gpu_batch(oc, batch["cls"][b:b + 1], batch["box"][b:b + 1], batch["cov"][b:b + 1], batch["anchors"],
                                batch["counts"][b:b + 1], emit_probs=False)
        for k in ("num_dets", "means", "covs", "cat_param", "cat_count", "nms_indices"):
            assert_bit_equal(getattr(res1, k)[0], getattr(res4, k)[b], k)


@pytest.mark.parametrize("depth", [2, 3, 8])
def test_pipelined_runs_equal_serial_runs(depth):
    """pipeline_depth >= 2: run i+1's head overlaps run i's tail on a second set of buffers.  A stream of
    different batches through one pipelined context gives the same bits as serial contexts, the last
    result is the one fetched, and device results of run i survive the issue of run i+1."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=45)
    B, rounds = 3, 11
    oc = oracle.OracleConfig()
    batches = [synthetic.to_numpy(synthetic.make_batch(spec, B, first_image_id=100 * i)) for i in range(rounds)]
    serial = [run_gpu_batch(oc, b["cls"], b["box"], b["cov"], b["anchors"], b["counts"], emit_probs=False)[1] for b in batches]
    N, A, K = batches[0]["cls"].shape[1:]
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc, pipeline_depth=depth))
    dev = [{k: torch.from_numpy(b[k]).cuda() for k in ("cls", "box", "cov", "anchors", "counts")} for b in batches]
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    keys = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx",
            "centre_scores")
    # (a) fetch after every run
    for i in range(rounds):
        d = dev[i]
        eng.run(d["cls"], d["box"], d["cov"], d["anchors"], d["counts"], stream=st.cuda_stream)
        res = eng.fetch()
        for k in keys:
            assert_bit_equal(getattr(res, k), getattr(serial[i], k), f"run {i}: {k}")
    # (b) back-to-back issue, only the last one fetched; twice to cover both lane parities
    for n in (rounds, rounds - 1):
        for i in range(n):
            d = dev[i]
            eng.run(d["cls"], d["box"], d["cov"], d["anchors"], d["counts"], stream=st.cuda_stream)
        res = eng.fetch()
        for k in keys:
            assert_bit_equal(getattr(res, k), getattr(serial[n - 1], k), f"back to back x{n}: {k}")
        r = oracle.run_image(oc, batches[n - 1]["cls"][1], batches[n - 1]["box"][1], batches[n - 1]["cov"][1],
                             batches[n - 1]["anchors"], batches[n - 1]["counts"][1])
        compare_image_with_oracle(eng, res, 1, r, K, check_probs=False)
    # (c) the synchronous host entry still works on a pipelined context
    b = batches[2]
    res = eng.run_host(b["cls"], b["box"], b["cov"], b["anchors"], b["counts"])
    for k in keys:
        assert_bit_equal(getattr(res, k), getattr(serial[2], k), f"host entry: {k}")


@pytest.mark.parametrize("hold", [False, True], ids=["caller_waits", "inputs_held"])
@pytest.mark.parametrize("graphs", ["2", "0"])
@pytest.mark.parametrize("depth", [1, 2, 4])
def test_streaming_retrieval_of_every_run(depth, graphs, hold, monkeypatch):
    """bod_fetch_async / bod_ticket_wait: a stream of different batches through one (pipelined) context, every
    run's result blocks copied out behind its own tail while later runs keep streaming; device results of run i
    survive the issue of runs i+1 .. i+L-1 (fetched oldest first after L back-to-back runs); tickets older than
    that are refused.  With graph replay (forced: BOD_GRAPHS=2, short runs default to stream launches) and with
    plain stream launches."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    from bayes_od_rc_b200._cabi import BodError
    monkeypatch.setenv("BOD_GRAPHS", graphs)
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=46)
    B, runs = 2, 13
    oc = oracle.OracleConfig()
    batches = [synthetic.to_numpy(synthetic.make_batch(spec, B, first_image_id=50 * i)) for i in range(runs)]
    serial = [run_gpu_batch(oc, b["cls"], b["box"], b["cov"], b["anchors"], b["counts"], emit_probs=False)[1] for b in batches]
    N, A, K = batches[0]["cls"].shape[1:]
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc, pipeline_depth=depth))
    if hold:        # every run has its own input tensors below and they live to the end of the test: heads of short runs
        eng.set_input_hold(True)   # alternate between two streams and the caller's stream is left alone (bod_set_input_hold)
    dev = [{k: torch.from_numpy(b[k]).cuda() for k in ("cls", "box", "cov", "anchors", "counts")} for b in batches]
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    keys = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx",
            "centre_scores")

    def check(res, i, what):
        for k in keys:
            assert_bit_equal(getattr(res, k), getattr(serial[i], k), f"{what}, run {i}: {k}")

    # (a) fetch_async right behind every run, collected `depth - 1` runs later
    tickets = []
    for i in range(runs):
        d = dev[i]
        eng.run(d["cls"], d["box"], d["cov"], d["anchors"], d["counts"], stream=st.cuda_stream)
        tickets.append(eng.fetch_async())
        assert tickets[-1] == eng.last_ticket
        j = i - (depth - 1)
        if j >= 0:
            check(eng.collect(tickets[j]), j, "streamed")
    for j in range(max(0, runs - depth + 1), runs):
        check(eng.collect(tickets[j]), j, "drained")
    # (b) L runs back to back, then every one of them fetched, oldest first: results survive in their lanes
    first = 3
    tk = []
    for i in range(first, first + depth):
        d = dev[i]
        eng.run(d["cls"], d["box"], d["cov"], d["anchors"], d["counts"], stream=st.cuda_stream)
        tk.append(eng.last_ticket)
    for n, t in enumerate(tk):
        eng.fetch_async(t)
        check(eng.collect(t), first + n, "kept in its lane")
    # (c) a ticket whose lane has been reused is refused
    with pytest.raises(BodError) as e:
        eng.fetch_async(tk[0] - 1 if depth > 1 else tk[0] - 1)
    assert e.value.status == _status("BOD_ERR_STATE")
    # (d) the input tensors of a lane change from run to run (graph nodes are re-pointed): covered by (a) --
    # and a run with the sampler instead of injected counts re-captures
    d = dev[0]
    eng.run(d["cls"], d["box"], d["cov"], d["anchors"], None, stream=st.cuda_stream)
    t = eng.fetch_async()
    res = eng.collect(t)
    _, ref = run_gpu_batch(oc, batches[0]["cls"], batches[0]["box"], batches[0]["cov"], batches[0]["anchors"], None,
                           emit_probs=False)
    for k in keys:
        assert_bit_equal(getattr(res, k), getattr(ref, k), f"sampler run after injected runs: {k}")


def _status(name):
    from bayes_od_rc_b200 import _cabi
    return getattr(_cabi, name)


@pytest.mark.parametrize("depth", [2, 3])
def test_pipelined_prefilter_every_run(depth):
    """Pipelined context with the pre-NMS filter (K2 stays on the head stream there): back-to-back runs of
    different batches, every run's results checked -- the lane's survivor counts must not be rewritten by the
    next run on that lane while its soft-NMS / fusion kernels still read them."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=57)
    B, runs = 2, 9
    oc = oracle.OracleConfig(pre_nms_top_k=150)
    batches = [synthetic.to_numpy(synthetic.make_batch(spec, B, first_image_id=70 * i)) for i in range(runs)]
    serial = [run_gpu_batch(oc, b["cls"], b["box"], b["cov"], b["anchors"], b["counts"], emit_probs=False)[1] for b in batches]
    assert len({int(s.num_survivors[0]) for s in serial}) >= 1
    N, A, K = batches[0]["cls"].shape[1:]
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc, pipeline_depth=depth))
    dev = [{k: torch.from_numpy(b[k]).cuda() for k in ("cls", "box", "cov", "anchors", "counts")} for b in batches]
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    tickets = []
    for i in range(runs):
        d = dev[i]
        eng.run(d["cls"], d["box"], d["cov"], d["anchors"], d["counts"], stream=st.cuda_stream)
        tickets.append(eng.fetch_async())
        j = i - (depth - 1)
        if j >= 0:
            res = eng.collect(tickets[j])
            for k in ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices"):
                assert_bit_equal(getattr(res, k), getattr(serial[j], k), f"run {j}: {k}")


def test_dlpack_inputs():
    """The TF-interop path of INTEGRATION.md: CUDA DLPack capsules and objects with __dlpack__ go through
    bayes_od_inference / the engine without a copy; host tensors that only speak DLPack are staged."""
    import torch
    from torch.utils.dlpack import to_dlpack
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    spec = synthetic.SceneSpec(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=47)
    B = 2
    nb = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig()
    _, ref = run_gpu_batch(oc, nb["cls"], nb["box"], nb["cov"], nb["anchors"], nb["counts"], emit_probs=False)
    N, A, K = nb["cls"].shape[1:]
    dev = {k: torch.from_numpy(nb[k]).cuda() for k in ("cls", "box", "cov", "anchors", "counts")}

    class OnlyDLPack:                       # what a foreign framework's tensor looks like: no data_ptr, no CAI
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None, **kw):
            return to_dlpack(self._t)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    keys = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices")
    for wrap in (to_dlpack, OnlyDLPack):
        eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc))
        eng.run(*[wrap(dev[k]) for k in ("cls", "box", "cov", "anchors", "counts")])
        res = eng.fetch()
        for k in keys:
            assert_bit_equal(getattr(res, k), getattr(ref, k), f"{wrap.__name__}: {k}")
    # through the drop-in, one image: DLPack-only objects for the head outputs (device) and the anchors (host)
    from bayes_od_rc_b200 import inference_utils as iu

    class HostDLPack(OnlyDLPack):
        def __init__(self, t):
            super().__init__(t)
            self.shape = tuple(t.shape)

    class DevDLPack(OnlyDLPack):
        def __init__(self, t):
            super().__init__(t)
            self.shape = tuple(t.shape)

    pred = {iu.ANCHORS_CLASS_PREDICTIONS_KEY: DevDLPack(dev["cls"][0]), iu.ANCHORS_BOX_PREDICTIONS_KEY: DevDLPack(dev["box"][0]),
            iu.ANCHORS_COVAR_PREDICTIONS_KEY: DevDLPack(dev["cov"][0])}
    model = lambda image, train_val_test="testing": pred      # noqa: E731
    sample_dict = {iu.IMAGE_NORMALIZED_KEY: np.zeros((1, spec.im_h, spec.im_w, 3), np.float32),
                   iu.ANCHORS_KEY: HostDLPack(torch.from_numpy(nb["anchors"][None].copy())),
                   iu.ORIGINAL_IM_SIZE_KEY: np.asarray([[spec.im_h, spec.im_w, 3]], np.int32)}
    bcfg = dict(dirichlet_prior=dict(type="non_informative"),
                gaussian_prior=dict(type="isotropic", isotropic_variance=100000.0), ranking_method="score")
    ncfg = dict(max_output_size=100, iou_threshold=0.5, soft_nms_sigma=0.5)
    out = iu.bayes_od_inference(model, sample_dict, bcfg, ncfg, use_full_covar=True, dataset_name="bdd", counts=nb["counts"][0])
    fused = iu.bayes_od_clustering(*[o.numpy() for o in out], affinity_threshold=0.5)
    d = int(ref.num_dets[0])
    assert_bit_equal(np.asarray(fused[1])[:, :, 0], ref.means[0, :d], "drop-in fused means from DLPack inputs")
    # consecutive calls without image_id use consecutive Philox streams on ONE cached engine
    n_eng = len(iu._engines)
    a = iu.bayes_od_inference(model, sample_dict, bcfg, ncfg, use_full_covar=True, dataset_name="bdd")
    b = iu.bayes_od_inference(model, sample_dict, bcfg, ncfg, use_full_covar=True, dataset_name="bdd")
    c = iu.bayes_od_inference(model, sample_dict, bcfg, ncfg, use_full_covar=True, dataset_name="bdd", image_id=12345)
    c2 = iu.bayes_od_inference(model, sample_dict, bcfg, ncfg, use_full_covar=True, dataset_name="bdd", image_id=12345)
    assert len(iu._engines) == n_eng, "a context per image id"
    assert not np.array_equal(np.asarray(a[0]), np.asarray(b[0])), "every image drew the same random stream"
    assert np.array_equal(np.asarray(c[0]), np.asarray(c2[0]))


def test_philox_sampler_matches_restatement():
    """Sampler mode: the counts the kernel draws equal the oracle's Philox restatement
    run on the kernel's own mean probabilities, bit for bit; every row sums to T."""
    spec = synthetic.SceneSpec(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=42)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 2, with_counts=False))
    oc = oracle.OracleConfig(seed=99, image_id_base=7)
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], None)
    for b in range(2):
        p = eng.probs(b)
        c = eng.sampled_counts(b)
        assert (c.sum(1) == 30).all()
        ref = oracle.philox_counts(p, 30, 99, 7 + b)
        assert_bit_equal(c, ref, "philox counts")
        # and the rest of the path on those counts is bit-exact again
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], batch["cov"][b], batch["anchors"], c)
        compare_image_with_oracle(eng, res, b, r, spec.K)


@pytest.mark.parametrize("K,N", [(11, 10), (8, 6), (4, 20), (7, 5)])
def test_softmax_rows_far_apart(K, N):
    """Mean class probabilities on logit rows the fast shift cannot handle: the moments kernel shifts every row by
    its background logit first and redoes rows whose sum leaves [1e-30, 1e30] with the row maximum (overflow,
    total underflow).  Rows: foreground far above / below the background, huge common offsets, one dominant column.
    Covers the unrolled (K = 4, 8, 11 with N a multiple of the ring depth) and the generic sample loop."""
    spec = synthetic.SceneSpec(im_h=96, im_w=160, N=N, K=K, g_min=3, g_max=5, box_hi=90., config_id=77)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 2, with_counts=False))
    cls = batch["cls"].copy()
    A = cls.shape[2]
    rng = np.random.default_rng(5)
    rows = rng.permutation(A)[:600]
    for i, a in enumerate(rows):
        kind = i % 6
        if kind == 0: cls[:, :, a, rng.integers(0, K - 1)] += 95.0            # exp overflows against the background shift
        elif kind == 1: cls[:, :, a, K - 1] -= 120.0                          # every foreground column overflows
        elif kind == 2: cls[:, :, a, :K - 1] -= 110.0                         # foreground underflows: background only
        elif kind == 3: cls[:, :, a, :] += 3000.0 * (1 if i % 12 == 3 else -1)   # common offset
        elif kind == 4: cls[:, rng.integers(0, N), a, rng.integers(0, K)] += 200.0   # one sample only
        else: cls[:, :, a, :] *= 40.0                                          # spread rows
    oc = oracle.OracleConfig(seed=3, image_id_base=11)
    eng, res = run_gpu_batch(oc, cls, batch["box"], batch["cov"], batch["anchors"], None)
    for b in range(2):
        p = eng.probs(b)
        ref = oracle.softmax_mean(cls[b], real="f64")
        assert np.isfinite(p).all()
        err = np.abs(p.astype(np.float64) - ref).max()
        assert err <= 2e-6, f"image {b}: mean class probabilities differ by {err}"
        c = eng.sampled_counts(b)
        assert_bit_equal(c, oracle.philox_counts(p, 30, 3, 11 + b), "philox counts")


@pytest.mark.parametrize("hw", [(320, 512), (384, 640)], ids=["pipeline_kernel", "unaligned_fallback_kernel"])
def test_moments_launch_clock_agrees_with_cuda_events(hw):
    """Pipelined contexts time the moments kernel with its own launch clock (%globaltimer stamps of the first CTA's start
    and the last CTA's end) instead of CUDA events; on a serial context both are available and have to agree.
    320x512 has 30 708 anchors (rows 16-byte aligned: the bulk-copy pipeline kernel), 384x640 has 46 035 (the fallback)."""
    spec = synthetic.SceneSpec(im_h=hw[0], im_w=hw[1], N=10, K=11, g_min=4, g_max=6, config_id=5)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 4, with_counts=False))
    oc = oracle.OracleConfig(seed=1)
    eng, _ = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], None, emit_probs=False)
    import torch
    dev = [torch.from_numpy(batch[k]).cuda() for k in ("cls", "box", "cov", "anchors")]
    eng.stage_ms_accum(); eng.moments_clock_accum()
    for _ in range(12):
        eng.run(*dev, None)
    ev, n_ev = eng.stage_ms_accum()
    clk, n_clk = eng.moments_clock_accum()
    assert n_ev == 12 and n_clk == 12
    ev_ms, clk_ms = ev["moments_filter"] / n_ev, clk / n_clk
    assert clk_ms > 0 and abs(ev_ms - clk_ms) <= 0.15 * ev_ms + 0.004, (ev_ms, clk_ms)   # events bracket launch latency too


def test_sampler_fast_path_equals_full_counts():
    """Without emit_probs the sampler skips the per-class draws of background-majority anchors;
    the survivors and everything downstream must not change."""
    spec = synthetic.SceneSpec(im_h=192, im_w=320, N=10, K=11, g_min=6, g_max=10, box_hi=150., config_id=44)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 3, with_counts=False))
    oc = oracle.OracleConfig(seed=7)
    _, res_full = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], None, emit_probs=True)
    _, res_fast = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], None, emit_probs=False)
    assert res_full.num_survivors.min() > 50
    for k in ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx",
              "centre_scores"):
        assert_bit_equal(getattr(res_fast, k), getattr(res_full, k), k)


def test_generated_anchors_bit_exact():
    import torch
    from bayes_od_rc_b200 import _cabi
    lib = _cabi.load()
    for (h, w) in [(720, 1280), (512, 1696), (375, 1242), (70, 90)]:
        A = lib.bod_generate_anchors(h, w, None, None)
        assert A == anchors_mod.num_anchors(h, w)
        buf = torch.empty(A, 4, device="cuda")
        assert lib.bod_generate_anchors(h, w, buf.data_ptr(), None) == A
        torch.cuda.synchronize()
        assert_bit_equal(buf.cpu().numpy(), oracle.generate_anchors(h, w), f"anchors {h}x{w}")
    # and against the digests of the reference generator's own output (tests/golden/make_anchor_digests.py)
    import hashlib
    import json
    import os
    digests = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "anchor_digests.json")))
    for key, want in digests.items():
        h, w = (int(v) for v in key.split("x"))
        buf = torch.empty(want["A"], 4, device="cuda")
        assert lib.bod_generate_anchors(h, w, buf.data_ptr(), None) == want["A"]
        torch.cuda.synchronize()
        assert hashlib.sha256(buf.cpu().numpy().tobytes()).hexdigest() == want["sha256"], key


def test_anchor_generate_mode_equals_tensor_mode():
    from bayes_od_rc_b200 import _cabi
    spec = synthetic.SceneSpec(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=43)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 2))
    oc = oracle.OracleConfig()
    _, res_t = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    _, res_g = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], None, batch["counts"], emit_probs=False,
                             anchor_mode=_cabi.ANCHORS_GENERATE, im_h=96, im_w=160)
    for k in ("num_dets", "means", "covs", "cat_param", "cat_count", "nms_indices"):
        assert_bit_equal(getattr(res_g, k), getattr(res_t, k), k)


def test_cluster_host_standalone():
    """bayes_od_clustering as its own entry point (bod_cluster_host) on a golden posterior."""
    from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
    g = load_golden("bdd_covar_k8")
    S, K = g["cnt_post"].shape
    mu = g["mu_post"][:, :, 0]
    corners = np.stack([mu[:, 0] - mu[:, 2] / 2, mu[:, 1] - mu[:, 3] / 2, mu[:, 0] + mu[:, 2] / 2, mu[:, 1] + mu[:, 3] / 2], 1)
    iou = oracle.iou_matrix(corners.astype(np.float32))
    eng = BayesODEngine(1, 10, max(S, 64), K, BayesODConfig(use_full_covar=True))
    fs, fm, fc, fn = eng.cluster_host(g["cnt_post"], mu, g["sig_post"], g["nms_indices"], iou, 0.5)
    mask = oracle.membership(corners.astype(np.float32), g["nms_indices"], 0.5)
    os_, om, oc_, on, _, _ = oracle.clustering(g["cnt_post"], mu, g["sig_post"], g["nms_indices"], mask, 70.0)
    assert fm.shape == (len(g["nms_indices"]), 4, 1) and fc.shape[1:] == (4, 4)
    assert_bit_equal(fm[:, :, 0], om, "means"); assert_bit_equal(fc, oc_, "covs")
    assert_bit_equal(fs, os_, "scores"); assert_bit_equal(fn, on, "counts")


def test_full_size_bdd_image_bit_exact():
    """One full BDD-shape image (A = 172 980, N = 10, K = 8) end to end."""
    spec = synthetic.SceneSpec(config_id=3)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 2))
    assert batch["cls"].shape[2] == 172980
    oc = oracle.OracleConfig()
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    for b in range(2):
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], batch["cov"][b], batch["anchors"], batch["counts"][b],
                             with_probs=False)
        assert 1000 < len(r.keep) < 10000
        compare_image_with_oracle(eng, res, b, r, 8, check_probs=False)


@pytest.mark.parametrize("case", ["covar_k8", "packed_k11_topk", "kendall_pipelined"])
def test_per_level_inputs_equal_concatenated(case):
    """bod_run_levels (head outputs still split per FPN level, before retinanet_model.py:89-112 concatenates
    them) gives the bits of bod_run on the concatenated tensors; the same context also still takes the
    concatenated form."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    kw = dict(covar_k8=dict(im_h=192, im_w=320, N=10, K=8, g_min=6, g_max=10, box_hi=150., config_id=81),
              packed_k11_topk=dict(im_h=100, im_w=180, N=6, K=11, g_min=6, g_max=10, box_hi=90., config_id=82, packed_cov=True),
              kendall_pipelined=dict(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=83))[case]
    okw = dict(covar_k8={}, packed_k11_topk=dict(cov_layout=2, pre_nms_top_k=150), kendall_pipelined=dict(use_full_covar=False))[case]
    depth = 3 if case == "kendall_pipelined" else 1
    spec = synthetic.SceneSpec(**kw)
    B = 3
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig(**okw)
    _, ref = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    la = anchors_mod.level_anchor_counts(spec.im_h, spec.im_w)
    N, A, K = batch["cls"].shape[1:]
    assert sum(la) == A and any(a % 128 for a in la)          # level boundaries do not fall on tile boundaries
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc, level_anchors=tuple(la), pipeline_depth=depth))
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()   # noqa: E731
    cuts = np.cumsum([0] + la)
    split = lambda x: [dev(x[:, :, cuts[l]:cuts[l + 1]]) for l in range(len(la))]   # noqa: E731
    cls_l, box_l, cov_l = split(batch["cls"]), split(batch["box"]), split(batch["cov"])
    anc, cnt = dev(batch["anchors"]), dev(batch["counts"])
    torch.cuda.synchronize()
    keys = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx",
            "centre_scores")
    for _ in range(depth + 1):
        eng.run_levels(cls_l, box_l, cov_l, anc, cnt)
    res = eng.fetch()
    for k in keys:
        assert_bit_equal(getattr(res, k), getattr(ref, k), f"per level: {k}")
    r = oracle.run_image(oc, batch["cls"][1], batch["box"][1], batch["cov"][1], batch["anchors"], batch["counts"][1])
    compare_image_with_oracle(eng, res, 1, r, K, check_probs=False)
    eng.run(dev(batch["cls"]), dev(batch["box"]), dev(batch["cov"]), anc, cnt)      # concatenated, same context
    res2 = eng.fetch()
    for k in keys:
        assert_bit_equal(getattr(res2, k), getattr(ref, k), f"concatenated on a level context: {k}")


@pytest.mark.parametrize("seed", range(40))
def test_random_configurations_bit_exact(seed):
    """Randomised scenes and knobs (class count, MC samples, soft / hard NMS, sigma down to values that
    drive scores into the denormals, thresholds, output sizes 1..255, ranking, priors, covariance layouts,
    dense and sparse scenes): every stage bit-exact against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.choice([2, 3, 4, 8, 11, 13]))
    N = int(rng.choice([2, 3, 6, 10]))
    dense = bool(rng.random() < 0.4)
    im_h, im_w = (int(rng.integers(64, 130)), int(rng.integers(80, 200)))
    spec_kw = dict(im_h=im_h, im_w=im_w, N=N, K=K, config_id=900 + seed, box_hi=float(min(im_h, im_w) - 4),
                   stray_frac=float(rng.choice([0.0, 0.002, 0.05])), packed_cov=bool(rng.random() < 0.3))
    if dense:
        spec_kw.update(g_min=1, g_max=2, box_lo=float(min(im_h, im_w) * 0.6), fg_iou=0.25)
    else:
        spec_kw.update(g_min=2, g_max=8, fg_iou=float(rng.choice([0.3, 0.4, 0.5])))
    soft = bool(rng.random() < 0.75)
    oc_kw = dict(soft_nms_sigma=float(rng.choice([0.02, 0.1, 0.5, 1.5])) if soft else 0.0,
                 iou_threshold=float(rng.choice([0.0, 0.3, 0.5, 0.7, 1.0])),
                 max_output_size=int(rng.choice([1, 7, 100, 255])),
                 use_full_covar=bool(rng.random() < 0.6),
                 cov_layout=2 if spec_kw["packed_cov"] else int(rng.choice([0, 1])),
                 dirichlet_prior=str(rng.choice(["non_informative", "None"])),
                 gaussian_prior=str(rng.choice(["isotropic", "isotropic", "None"])),
                 ranking_method=str(rng.choice(["score", "score", "joint_entropy"])))
    if oc_kw["cov_layout"] == 0 and oc_kw["gaussian_prior"] == "isotropic" and N < 6:
        # no covariance head and N <= 4 samples: the likelihood covariance is the rank-deficient sample
        # covariance, its inverse (inference_utils.py:101) is not finite and every box is NaN in the reference
        # too -- outside the parity contract (min / max of NaN is implementation-defined)
        N = 6
        spec_kw["N"] = N
    if rng.random() < 0.3:
        oc_kw.update(pre_nms_top_k=int(rng.integers(5, 200)))
    if rng.random() < 0.2:
        oc_kw.update(score_threshold=float(rng.choice([0.2, 0.5])))
    spec = synthetic.SceneSpec(**spec_kw)
    B = int(rng.integers(1, 4))
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B))
    oc = oracle.OracleConfig(**oc_kw)
    cov = batch["cov"] if oc.cov_layout else None
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], cov, batch["anchors"], batch["counts"],
                             pipeline_depth=int(rng.choice([1, 1, 3])))
    for b in range(B):
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], None if cov is None else cov[b], batch["anchors"],
                             batch["counts"][b])
        compare_image_with_oracle(eng, res, b, r, K)


@pytest.mark.parametrize("case", [
    dict(soft_nms_sigma=0.1, max_output_size=255), dict(soft_nms_sigma=1.5, max_output_size=200, iou_threshold=0.3),
    dict(soft_nms_sigma=0.0, iou_threshold=0.7, max_output_size=255), dict(soft_nms_sigma=0.0, iou_threshold=0.3),
    dict(soft_nms_sigma=0.5, max_output_size=255, ranking_method="joint_entropy"),
    dict(soft_nms_sigma=0.25, max_output_size=150, use_full_covar=False, pre_nms_top_k=2000)])
def test_full_size_image_other_knobs(case):
    """Full BDD-shape images (S ~ 3000) with NMS settings away from the YAML defaults: long pending lists
    (small sigma, 255 centres), hard NMS, joint-entropy ranking, pre-NMS top-k."""
    spec = synthetic.SceneSpec(config_id=60 + len(case), K=8)
    batch = synthetic.to_numpy(synthetic.make_batch(spec, 2))
    oc = oracle.OracleConfig(**case)
    eng, res = run_gpu_batch(oc, batch["cls"], batch["box"], batch["cov"], batch["anchors"], batch["counts"], emit_probs=False)
    for b in range(2):
        r = oracle.run_image(oc, batch["cls"][b], batch["box"][b], batch["cov"][b], batch["anchors"], batch["counts"][b],
                             with_probs=False)
        assert len(r.keep) > 1000
        compare_image_with_oracle(eng, res, b, r, 8, check_probs=False)


@pytest.mark.parametrize("seed", range(6))
def test_validation_random(seed):
    rng = np.random.default_rng(500 + seed)
    K = int(rng.choice([2, 4, 8, 11]))
    spec = synthetic.SceneSpec(im_h=int(rng.integers(64, 200)), im_w=int(rng.integers(96, 330)), N=2, K=K, g_min=2, g_max=9,
                               box_hi=60., config_id=950 + seed, stray_frac=float(rng.choice([0.0, 0.01])))
    B = int(rng.integers(1, 4))
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B, with_counts=False))
    cls, box = batch["cls"][:, 0], batch["box"][:, 0]
    ckw = dict(soft_nms_sigma=float(rng.choice([0.0, 0.1, 0.5])), iou_threshold=float(rng.choice([0.3, 0.5])),
               max_output_size=int(rng.choice([5, 100, 255])))
    mode = int(rng.integers(0, 3))
    scaling = None if mode == 0 else (mode, (4., 2., 4., 2.), (float(spec.im_h - 8), float(spec.im_w - 4)), (375., 1242.))
    okw = dict(ckw, scale_mode=mode)
    if mode:
        okw.update(shift=scaling[1], norm_hw=scaling[2], scale_hw=scaling[3])
    eng, res = _run_validate(cls, box, batch["anchors"], scaling=scaling, **ckw)
    for b in range(B):
        _compare_validate(eng, res, b, oracle.val_postprocess(cls[b], box[b], batch["anchors"], **okw), K)


FULL_BATCHES = {
    # BASELINE.json configurations at full size, every image of the batch bit for bit against the oracle
    # name: (SceneSpec kwargs, OracleConfig kwargs, B, pipeline_depth)
    "bdd_covar_b32_k11": (dict(N=10, K=11, config_id=3), dict(), 32, 1),                      # the benchmarked configuration
    "bdd_covar_b32_k11_pipelined": (dict(N=10, K=11, config_id=3), dict(), 32, 4),
    "kitti_512x1696_b8_n20_k4": (dict(im_h=512, im_w=1696, N=20, K=4, config_id=4),
                                 dict(scale_v=375 / 512, scale_u=1242 / 1696), 8, 1),
    "kitti_raw_375x1242_b8_n20_k4": (dict(im_h=375, im_w=1242, N=20, K=4, config_id=6), dict(), 8, 1),
    "bdd_kendall_b8_k8": (dict(N=10, K=8, config_id=2), dict(use_full_covar=False), 8, 1),
    "stress_b4_n40_k11_topk": (dict(N=40, K=11, config_id=5, g_min=80, g_max=120, fg_iou=0.2, fg_logit=1.0, bg_logit_for_fg=0.0,
                                    stray_frac=0.02), dict(score_threshold=0.01, pre_nms_top_k=10000), 4, 1),
}


@pytest.mark.parametrize("sampler", ["injected", "philox"])
@pytest.mark.parametrize("name", sorted(FULL_BATCHES))
def test_full_size_batches_bit_exact(name, sampler):
    """Full-size batches of BASELINE.json's configurations (the benchmarked B = 32, K = 11, full covariance among
    them): every image's padded result block equals the oracle's, with injected counts and with the in-kernel
    Philox sampler (its counts are checked against the restatement on the kernel's own mean probabilities, the
    oracle then runs on those counts).  Inputs are generated on the device by the seeded generator."""
    import torch
    from gpu_common import engine_config_from_oracle
    from bayes_od_rc_b200.engine import BayesODEngine
    spec_kw, oc_kw, B, depth = FULL_BATCHES[name]
    if sampler == "philox" and depth > 1:
        pytest.skip("one sampler case per shape")
    spec = synthetic.SceneSpec(**spec_kw)
    batch = synthetic.make_batch(spec, B, device="cuda", with_counts=(sampler == "injected"), first_image_id=40)
    N, A, K = batch["cls"].shape[1:]
    oc = oracle.OracleConfig(seed=4321, image_id_base=40, **oc_kw)
    cap = min(A, 49152)
    eng = BayesODEngine(B, N, A, K, engine_config_from_oracle(oc, pipeline_depth=depth, emit_probs=(sampler == "philox"),
                                                              max_survivors=cap))
    counts = batch["counts"] if sampler == "injected" else None
    for _ in range(3 if depth > 1 else 1):                  # pipelined: stream path, then graph replays
        eng.run(batch["cls"], batch["box"], batch["cov"], batch["anchors"], counts)
    res = eng.fetch()
    if sampler == "philox":
        cnt = np.stack([eng.sampled_counts(b) for b in range(B)])
        for b in range(0, B, max(1, B // 4)):               # the restatement is slow in Python: a quarter of the images
            assert_bit_equal(cnt[b], oracle.philox_counts(eng.probs(b), 30, 4321, 40 + b), f"image {b}: philox counts")
    else:
        cnt = batch["counts"].cpu().numpy()
    h = {k: batch[k].cpu().numpy() for k in ("cls", "box", "cov", "anchors")}
    ref = oracle.run_batch(oc, h["cls"], h["box"], h["cov"], h["anchors"], cnt, nthreads=16)
    assert ref["num_survivors"].min() > 1000
    if "topk" in name:
        assert (ref["num_survivors"] == 10000).all()
    for k in ("num_survivors", "num_dets", "nms_indices", "centre_anchor_idx", "means", "cat_param", "cat_count"):
        assert_bit_equal(getattr(res, k), ref[k], f"{name}: {k}")
    assert_bit_equal(res.covs.reshape(B, -1, 16), ref["covs"], f"{name}: covs")


def test_full_size_batch_properties():
    """The bench workload at full size (8 BDD-shape images, N = 10, K = 11, Philox sampler) through
    size-independent properties: survivors ascending, centres unique and in selection-score order, every
    centre a member of its own cluster, Dirichlet count invariants (rows sum to
    31 m for m <= 3 members, 93 otherwise), symmetric positive fused covariances, identical bits from a
    second run, from a pipelined context and from a different sharding of the same global images."""
    import torch
    from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
    B, K = 8, 11
    spec = synthetic.SceneSpec(N=10, K=K, config_id=3)
    batch = synthetic.make_batch(spec, B, device="cuda", with_counts=False, first_image_id=16)
    A = batch["anchors"].shape[0]
    assert A == 172980

    def run(b0, nb, depth=1, repeat=1):
        cfg = BayesODConfig(use_full_covar=True, seed=1234, image_id_base=16 + b0, max_survivors=32768, pipeline_depth=depth)
        eng = BayesODEngine(nb, 10, A, K, cfg)
        for _ in range(repeat):
            eng.run(batch["cls"][b0:b0 + nb], batch["box"][b0:b0 + nb], batch["cov"][b0:b0 + nb], batch["anchors"], None)
        return eng, eng.fetch()

    eng, res = run(0, B)
    keys = ("num_dets", "num_survivors", "means", "covs", "cat_param", "cat_count", "nms_indices", "centre_anchor_idx",
            "centre_scores")
    for b in range(B):
        S, D = int(res.num_survivors[b]), int(res.num_dets[b])
        assert 1000 < S < 10000 and D == 100
        sv = eng.survivors(b)
        assert (np.diff(sv["anchor_idx"]) > 0).all()
        assert np.allclose(sv["counts"].sum(1), 31.0, atol=1e-3)
        idx = res.nms_indices[b, :D]
        assert len(np.unique(idx)) == D and idx.min() >= 0 and idx.max() < S
        assert (np.diff(res.centre_scores[b, :D]) <= 0).all()
        assert_bit_equal(res.centre_anchor_idx[b, :D], sv["anchor_idx"][idx], "centre anchors")
        mem = oracle.mask_to_bool(eng.members(b, S, D), S)               # [D,S]
        assert mem[np.arange(D), idx].all()
        m = mem.sum(1)
        tot = res.cat_count[b, :D].sum(1)
        assert np.allclose(tot, np.where(m > 3, 93.0, 31.0 * m), atol=1e-2)
        assert np.allclose(res.cat_param[b, :D].sum(1), 1.0, atol=1e-4)
        cov = res.covs[b, :D].astype(np.float64)
        assert np.allclose(cov, cov.transpose(0, 2, 1), rtol=1e-3, atol=1e-3)
        assert (np.linalg.eigvalsh((cov + cov.transpose(0, 2, 1)) / 2) > 0).all()
    _, again = run(0, B, repeat=3)
    _, piped = run(0, B, depth=3, repeat=4)
    for k in keys:
        assert_bit_equal(getattr(again, k), getattr(res, k), f"second run: {k}")
        assert_bit_equal(getattr(piped, k), getattr(res, k), f"pipelined: {k}")
    _, shard = run(5, 3)                                  # images 5..7 as their own shard
    for k in keys:
        assert_bit_equal(getattr(shard, k), getattr(res, k)[5:8], f"resharded: {k}")


@pytest.mark.parametrize("name", ["bdd_covar_k8", "kitti_k4_n8", "no_survivor"])
def test_dropin_inference_utils(name):
    """The reference-facing pair (same names / arguments / return structure as
    inference_utils.py:13-217 and :285-364), driven the way run_inference.py:137-161 drives it."""
    import torch
    from bayes_od_rc_b200 import inference_utils as fast
    g = load_golden(name)
    meta = g["meta"]
    pred = {fast.ANCHORS_CLASS_PREDICTIONS_KEY: torch.from_numpy(g["cls"]).cuda(),
            fast.ANCHORS_BOX_PREDICTIONS_KEY: torch.from_numpy(g["box"]).cuda(),
            fast.ANCHORS_COVAR_PREDICTIONS_KEY: torch.from_numpy(g["cov"]).cuda()}
    model = lambda image, train_val_test="testing": pred      # noqa: E731
    h, w = meta["image_shape"]
    sample_dict = {fast.IMAGE_NORMALIZED_KEY: np.zeros((1, h, w, 3), np.float32), fast.ANCHORS_KEY: g["anchors"][None],
                   fast.ORIGINAL_IM_SIZE_KEY: np.asarray([[meta["orig_size"][0], meta["orig_size"][1], 3]], np.int32)}
    cfg = meta["cfg"]
    out = fast.bayes_od_inference(model, sample_dict, cfg["bayes_od_config"], cfg["nms_config"],
                                  use_full_covar=cfg["use_full_covar"], dataset_name=meta["dataset_name"],
                                  counts=g["counts"])
    counts, means, covs, nms_indices, iou_mat = [o.numpy() for o in out]          # run_inference.py:141-145
    S, D = len(g["cnt_post"]), len(g["nms_indices"])
    assert counts.shape == (S, g["cls"].shape[2]) and means.shape == (S, 4, 1) and covs.shape == (S, 4, 4)
    assert np.array_equal(nms_indices, g["nms_indices"])
    if means.size > 0:                                                               # run_inference.py:147
        assert np.array_equal(counts, g["cnt_post"])
        assert within_tol(means, g["mu_post"]).all()
        fs, fm, fc, fn = fast.bayes_od_clustering(counts, means, covs, nms_indices, iou_mat,
                                                  affinity_threshold=cfg["nms_config"]["iou_threshold"])
        assert fm.shape == (D, 4, 1) and fc.shape == (D, 4, 4) and fs.shape == fn.shape == (D, counts.shape[1])
        assert all(a.dtype == np.float32 for a in (fs, fm, fc, fn))
        assert within_tol(fm, g["final_means"]).all()
        assert np.squeeze(fm, axis=2).shape == (D, 4)                                # run_inference.py:151
        with pytest.raises(ValueError):
            fast.bayes_od_clustering(counts, means, covs, nms_indices, iou_mat, affinity_threshold=0.7)
    else:
        assert nms_indices.size == 0


# --------------------------------------------------------------------------
# validation post-process (validation_utils.post_process_predictions)
# --------------------------------------------------------------------------
def _run_validate(cls, box, anchors, scaling=None, **cfg_kw):
    """cls [B,A,K], box [B,A,4] numpy -> (engine, results)."""
    import torch
    from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
    B, A, K = cls.shape
    eng = BayesODEngine(B, 1, A, K, BayesODConfig(**cfg_kw))
    t = [torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda() for x in (cls, box, anchors)]
    torch.cuda.synchronize()
    eng.validate(*t, scaling=scaling, stream=torch.cuda.current_stream().cuda_stream)
    return eng, eng.fetch()


def _compare_validate(eng, res, b, r, K):
    S, D = len(r.keep), len(r.nms_indices)
    assert int(res.num_survivors[b]) == S and int(res.num_dets[b]) == D
    sv = eng.survivors(b)
    assert_bit_equal(sv["anchor_idx"], r.keep, "kept anchors")
    assert_bit_equal(sv["counts"], r.probs, "softmax rows")
    assert_bit_equal(sv["corners"], r.corners, "corners")
    assert_bit_equal(sv["scores"], r.scores, "top scores")
    assert_bit_equal(res.nms_indices[b, :D], r.nms_indices, "nms_indices")
    assert_bit_equal(res.centre_scores[b, :D], r.nms_scores, "scores at selection")
    assert_bit_equal(res.cat_param[b, :D], r.classes_out, "classes_out")
    assert_bit_equal(res.means[b, :D], r.corners_out, "corners_out")
    assert not res.cat_param[b, D:].any() and not res.means[b, D:].any() and not res.covs[b].any()


@pytest.mark.parametrize("name", val_golden_cases())
def test_validation_golden(name):
    g = load_golden(name)
    mode, shift, norm_hw, scale_hw = val_scaling_of(g["meta"])
    eng, res = _run_validate(g["cls"][None], g["box"][None], g["anchors"],
                             scaling=None if mode == 0 else (mode, shift, norm_hw, scale_hw))
    r = oracle.val_postprocess(g["cls"], g["box"], g["anchors"], scale_mode=mode, shift=shift, norm_hw=norm_hw, scale_hw=scale_hw)
    _compare_validate(eng, res, 0, r, g["cls"].shape[1])
    D = len(g["classes_out"])                      # and against the reference's own function
    assert int(res.num_dets[0]) == D
    if D:
        assert within_tol(res.cat_param[0, :D], g["classes_out"]).all()
        assert within_tol(res.means[0, :D], g["corners_out"], rtol=1e-5, atol=1e-4).all()


@pytest.mark.parametrize("case", ["bdd_k8", "kitti_k4", "coco_k11", "hard_nms"])
def test_validation_batch_bit_exact(case):
    kw = dict(bdd_k8=dict(im_h=192, im_w=320, N=2, K=8, g_min=6, g_max=10, box_hi=150., config_id=71),
              kitti_k4=dict(im_h=128, im_w=424, N=2, K=4, g_min=5, g_max=9, box_hi=120., config_id=72),
              coco_k11=dict(im_h=160, im_w=160, N=2, K=11, g_min=5, g_max=9, box_hi=120., config_id=73),
              hard_nms=dict(im_h=96, im_w=160, N=2, K=8, g_min=4, g_max=6, box_hi=90., config_id=74))[case]
    spec = synthetic.SceneSpec(**kw)
    B = 3
    batch = synthetic.to_numpy(synthetic.make_batch(spec, B, with_counts=False))
    cls, box = batch["cls"][:, 0], batch["box"][:, 0]
    scaling, okw, ckw = None, {}, {}
    if case == "kitti_k4":
        scaling = (1, (0, 0, 0, 0), (128., 424.), (94., 311.)); okw = dict(scale_mode=1, norm_hw=(128., 424.), scale_hw=(94., 311.))
    if case == "coco_k11":
        scaling = (2, (8., 0., 8., 0.), (144., 160.), (480., 533.))
        okw = dict(scale_mode=2, shift=(8., 0., 8., 0.), norm_hw=(144., 160.), scale_hw=(480., 533.))
    if case == "hard_nms":
        ckw = dict(soft_nms_sigma=0.0, max_output_size=50); okw = dict(soft_nms_sigma=0.0, max_output_size=50)
    eng, res = _run_validate(cls, box, batch["anchors"], scaling=scaling, **ckw)
    for b in range(B):
        r = oracle.val_postprocess(cls[b], box[b], batch["anchors"], **okw)
        assert len(r.keep) > 10
        _compare_validate(eng, res, b, r, spec.K)


@pytest.mark.parametrize("name", ["val_kitti_k4", "val_coco_k8", "val_bdd_k8"])
@pytest.mark.parametrize("host_anchors", ["numpy", "dlpack_cpu"])
def test_validation_dropin(name, host_anchors):
    """bayes_od_rc_b200.validation_utils.post_process_predictions driven the way run_validation.py:143-147 does:
    the kitti, coco (constants.IMAGE_PADDING_KEY = 'paddings_applied') and bdd branches; `anchors` from the host as
    numpy or as a CPU tensor that only speaks DLPack (a tf.data pipeline hands them over on the host)."""
    import torch
    from bayes_od_rc_b200 import validation_utils as fast
    assert fast.IMAGE_PADDING_KEY == "paddings_applied"                                 # src/core/constants.py:55
    g = load_golden(name)
    meta = g["meta"]
    pred = {fast.ANCHORS_CLASS_PREDICTIONS_KEY: torch.from_numpy(g["cls"][None]).cuda(),
            fast.ANCHORS_BOX_PREDICTIONS_KEY: torch.from_numpy(g["box"][None]).cuda()}
    h, w = meta["image_shape"]
    anchors = g["anchors"][None]
    if host_anchors == "dlpack_cpu":
        class OnlyDLPack:
            def __init__(self, t):
                self._t, self.shape = t, tuple(t.shape)

            def __dlpack__(self, stream=None, **kw):
                return self._t.__dlpack__()

            def __dlpack_device__(self):
                return self._t.__dlpack_device__()
        anchors = OnlyDLPack(torch.from_numpy(np.ascontiguousarray(anchors)))
    sample_dict = {fast.IMAGE_NORMALIZED_KEY: np.zeros((1, h, w, 3), np.float32), fast.ANCHORS_KEY: anchors,
                   fast.ORIGINAL_IM_SIZE_KEY: np.asarray([[meta["orig_size"][0], meta["orig_size"][1], 3]], np.int32)}
    if meta["dataset_name"] == "coco":
        sample_dict[fast.IMAGE_PADDING_KEY] = np.asarray([meta["padding"]], np.float32)
    output_classes, output_boxes = fast.post_process_predictions(sample_dict, pred, dataset_name=meta["dataset_name"])
    output_boxes = output_boxes.numpy(); output_classes = output_classes.numpy()       # run_validation.py:146-147
    assert output_classes.shape == g["classes_out"].shape and output_boxes.shape == g["corners_out"].shape
    assert within_tol(output_classes, g["classes_out"]).all()
    assert within_tol(output_boxes, g["corners_out"], rtol=1e-5, atol=1e-4).all()
