"""CPU tests: the C-ABI library loads and exports every symbol the header
declares, the host-side mirror maps the reference's config dicts correctly, the
product refuses to run without a GPU (no CPU fallback), and the multi-rank
helpers work over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "bayesod.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bod_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bayes_od_rc_b200 import _cabi, build
    build.build_library()
    lib = _cabi.load()
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bayesod.h but not exported"
        assert n in _cabi.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.bod_abi_version() == 8
    assert lib.bod_status_string(-5).decode().startswith("more survivors")


def test_config_struct_layout_matches_header():
    """sizeof/offsets of the ctypes mirror follow the C struct (checked by compiling a probe)."""
    from bayes_od_rc_b200._cabi import BodConfig
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "bayesod.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(bod_config), offsetof(bod_config, seed),
        offsetof(bod_config, image_id_base), offsetof(bod_config, anchor_mode), offsetof(bod_config, emit_probs),
        offsetof(bod_config, pipeline_depth), offsetof(bod_config, level_anchors), sizeof(bod_val_scaling)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")], check=True)
        out = subprocess.run([os.path.join(d, "p")], check=True, capture_output=True, text=True).stdout.split()
    from bayes_od_rc_b200._cabi import BodValScaling
    got = [ctypes.sizeof(BodConfig), BodConfig.seed.offset, BodConfig.image_id_base.offset, BodConfig.anchor_mode.offset,
           BodConfig.emit_probs.offset, BodConfig.pipeline_depth.offset, BodConfig.level_anchors.offset,
           ctypes.sizeof(BodValScaling)]
    assert got == [int(x) for x in out]


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayes_od_rc_b200._cabi import BodError
    from bayes_od_rc_b200.engine import BayesODEngine
    with pytest.raises(BodError) as e:
        BayesODEngine(1, 10, 1161, 8)
    assert "BOD_ERR_CUDA" in str(e.value)
    from bayes_od_rc_b200 import inference_utils as fast
    with pytest.raises(BodError):
        fast.bayes_od_clustering(np.ones((4, 8), np.float32), np.zeros((4, 4, 1), np.float32),
                                 np.tile(np.eye(4, dtype=np.float32), (4, 1, 1)), np.array([0]), np.ones((4, 4), np.float32), 0.5)


def test_create_rejects_bad_configs_before_touching_the_device():
    """Argument validation of bod_create (runs without a GPU: the checks come before cudaSetDevice)."""
    from bayes_od_rc_b200 import _cabi
    from bayes_od_rc_b200.engine import BayesODConfig
    lib = _cabi.load()

    def create(B=1, N=10, A=1161, K=8, **kw):
        cfg = BayesODConfig(**kw).to_c(B, N, A, K)
        ctx = ctypes.c_void_p()
        rc = lib.bod_create(ctypes.byref(ctx), 0, ctypes.byref(cfg))
        if rc == 0:
            lib.bod_destroy(ctx)
        return rc, (lib.bod_last_error(None) or b"").decode()

    for kw, what in [(dict(B=0), "B must"), (dict(N=0), "N (mc"), (dict(K=1), "unsupported K"), (dict(K=14), "unsupported K"), (dict(K=81), "COCO"),
                     (dict(max_output_size=0), "max_output_size"), (dict(max_output_size=256), "max_output_size"),
                     (dict(iou_threshold=-0.1), "iou_threshold"), (dict(soft_nms_sigma=-1.0), "soft_nms_sigma"),
                     (dict(num_draws=0), "num_draws"), (dict(pre_nms_top_k=-1), "pre_nms_top_k"),
                     (dict(level_anchors=(1000, 100)), "do not sum"), (dict(level_anchors=(1161, 0)), "positive"),
                     (dict(anchor_mode=_cabi.ANCHORS_GENERATE, im_h=64, im_w=64), "anchor_mode=GENERATE")]:
        pos = {k: kw.pop(k) for k in ("B", "N", "K") if k in kw}
        rc, msg = create(**pos, **kw)
        assert rc == -1 and what in msg, (kw, rc, msg)


def test_npy_writers_are_byte_identical_to_numpy(tmp_path):
    """bod_write_results_npy vs what run_inference.py:241-244 writes with np.save (no GPU needed)."""
    from bayes_od_rc_b200 import writers
    rng = np.random.default_rng(0)
    B, D, K = 5, 100, 11
    nd = np.array([100, 0, 7, 1, 63], np.int32)
    res = dict(num_dets=nd, means=rng.normal(size=(B, D, 4)).astype(np.float32),
               covs=rng.normal(size=(B, D, 4, 4)).astype(np.float32),
               cat_param=rng.random((B, D, K)).astype(np.float32), cat_count=rng.random((B, D, K)).astype(np.float32))
    ids = [f"img_{b:04d}" for b in range(B)]
    dirs = [str(tmp_path / n) for n in ("mean", "cov", "cat_param", "cat_count")]
    writers.save_batch(res, ids, *dirs, nthreads=3)
    for b in range(B):
        d = int(nd[b])
        empty = np.zeros((0, 4, 1), np.float32)
        want = [res["means"][b, :d], res["covs"][b, :d], res["cat_param"][b, :d], res["cat_count"][b, :d]] if d else [empty] * 4
        for dirname, arr in zip(dirs, want):
            ref = tmp_path / "ref.npy"
            np.save(ref, np.ascontiguousarray(arr))
            got = open(os.path.join(dirname, ids[b] + ".npy"), "rb").read()
            assert got == open(ref, "rb").read(), (b, dirname)
            assert np.array_equal(np.load(os.path.join(dirname, ids[b] + ".npy")), arr)
    from bayes_od_rc_b200._cabi import BodError
    with pytest.raises(BodError):
        import ctypes as C
        from bayes_od_rc_b200 import _cabi
        r = _cabi.BodHostResults(num_dets=nd.ctypes.data, means=res["means"].ctypes.data, covs=res["covs"].ctypes.data,
                                 cat_param=res["cat_param"].ctypes.data, cat_count=res["cat_count"].ctypes.data)
        idp = (C.c_char_p * B)(*[i.encode() for i in ids])
        rc = _cabi.load().bod_write_results_npy(C.byref(r), B, D, K, b"/nonexistent/a", b"/nonexistent/b", b"/nonexistent/c",
                                                b"/nonexistent/d", idp, 2)
        raise BodError(rc, "x") if rc else AssertionError("expected a failure")


def test_bdd_json_writer_matches_reference_bytes(tmp_path):
    """BddJsonWriter vs predictions_to_bdd_format + json.dump of the reference, executed verbatim when the
    fixture was minted (tests/golden/make_writer_golden.py): same bytes."""
    from bayes_od_rc_b200 import writers
    g = np.load(os.path.join(ROOT, "tests", "golden", "writers_bdd.npz"))
    path = tmp_path / "predictions.json"
    with writers.BddJsonWriter(path, [str(c) for c in g["categories"]]) as w:
        for i in range(int(g["n_blocks"])):
            w.append(dict(num_dets=g[f"num_dets{i}"], means=g[f"means{i}"], cat_param=g[f"cat_param{i}"]),
                     [str(x) for x in g[f"ids{i}"]])
    got = open(path, "rb").read()
    assert got == g["json"].tobytes()
    import json
    assert len(json.loads(got)) == 24
    with writers.BddJsonWriter(path, ["car"]):
        pass
    assert open(path, "rb").read() == g["json_empty"].tobytes() == b"[]"
    # a class block override (map_dataset_classes output) with fewer columns
    with writers.BddJsonWriter(path, ["car", "person"]) as w:
        w.append(dict(num_dets=g["num_dets0"], means=g["means0"], cat_param=g["cat_param0"]), [str(x) for x in g["ids0"]],
                 cat_param=np.ascontiguousarray(g["cat_param0"][:, :, :3]))
    assert all(e["category"] in ("car", "person") for e in json.loads(open(path).read()))
    from bayes_od_rc_b200._cabi import BodError
    with pytest.raises(BodError):
        writers.BddJsonWriter("/nonexistent/dir/predictions.json", ["car"])


def test_kitti_txt_writer_matches_reference_bytes(tmp_path):
    """save_kitti_txt_batch vs predictions_to_kitti_format + np.savetxt(fmt='%s', newline='\\r\\n') of the reference."""
    from bayes_od_rc_b200 import writers
    g = np.load(os.path.join(ROOT, "tests", "golden", "writers_kitti.npz"))
    ids = [str(x) for x in g["ids"]]
    writers.save_kitti_txt_batch(dict(num_dets=g["num_dets"], means=g["means"], cat_param=g["cat_param"]), ids,
                                 tmp_path / "data", nthreads=3)
    for b, i in enumerate(ids):
        assert open(tmp_path / "data" / (i + ".txt"), "rb").read() == g[f"txt{b}"].tobytes(), i
    assert len(g["txt0"]) > 0 and len(g["txt1"]) == 0


def test_float_formats_match_python_and_numpy():
    """The two number formats of the text writers against the interpreters themselves: float.__repr__ (json) and
    str(numpy.float32) (np.savetxt of the KITTI rows), on random bit patterns and the layout thresholds."""
    import ctypes as C
    from bayes_od_rc_b200 import _cabi
    lib = _cabi.load()
    buf = C.create_string_buffer(64)

    def fmt(v, style):
        n = lib.bod_format_float(float(v), style, buf, 64)
        assert n > 0
        return buf.value.decode()

    rng = np.random.default_rng(5)
    f32 = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    f32 = np.concatenate([f32[np.isfinite(f32)], np.float32([0.0, -0.0, 1e-4, 9.9999e-5, 1e16, 9.9999e15, 1.0, 100.0, 0.1,
                                                           16777216.0, 1e-45, 3.4028235e38, 5e-324, 1e15, 123456.79])])
    for v in f32:
        assert fmt(v, 1) == str(v), (v, fmt(v, 1))
        assert fmt(v, 0) == repr(float(v)), (v, fmt(v, 0))
    f64 = rng.integers(0, 2 ** 63, 20000, dtype=np.uint64).view(np.float64)
    f64 = np.concatenate([f64[np.isfinite(f64)], [1e16, 9999999999999998.0, 1e-4, 9.999999999999999e-05, 1e22, 1e23, 5e-324,
                                                  1.7976931348623157e308, 0.30000000000000004, 2.0 ** 53]])
    for v in f64:
        assert fmt(v, 0) == repr(float(v)), v
        assert fmt(-v, 0) == repr(float(-v)), v
    assert fmt(float("nan"), 0) == "NaN" and fmt(float("inf"), 0) == "Infinity" and fmt(float("-inf"), 0) == "-Infinity"
    assert fmt(float("nan"), 1) == "nan" and fmt(float("-inf"), 1) == "-inf"
    assert lib.bod_format_float(1.2345678901234567, 0, buf, 4) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayes_od_rc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "bayesod_oracle" not in txt, f


def test_config_from_reference_yaml_dicts():
    from bayes_od_rc_b200.engine import BayesODConfig
    bayes = dict(ranking_method='score', dirichlet_prior=dict(type='non_informative'),
                 gaussian_prior=dict(type='isotropic', isotropic_variance=100000.0), fusion_method='none')
    nms = dict(max_output_size=100, iou_threshold=0.5, soft_nms_sigma=0.5)       # retinanet_bdd.yaml:128-131
    c = BayesODConfig.from_reference(bayes, nms, use_full_covar=True).to_c(1, 10, 172980, 8)
    assert (c.B, c.N, c.A, c.K) == (1, 10, 172980, 8)
    assert c.use_full_covar == 1 and c.dirichlet_prior == 1 and c.gaussian_prior == 1 and c.ranking_method == 0
    assert c.max_output_size == 100 and c.iou_threshold == 0.5 and c.soft_nms_sigma == 0.5
    assert c.isotropic_variance == 100000.0 and c.cov_calibration == 70.0 and c.num_draws == 30
    bayes['dirichlet_prior']['type'] = 'None'; bayes['ranking_method'] = 'joint_entropy'
    c = BayesODConfig.from_reference(bayes, nms).to_c(1, 10, 100, 8)
    assert c.dirichlet_prior == 0 and c.ranking_method == 1 and c.use_full_covar == 0


def test_device_ptr_rejects_host_memory():
    import torch
    from bayes_od_rc_b200.engine import device_ptr
    with pytest.raises(TypeError):
        device_ptr(torch.zeros(4))
    with pytest.raises(TypeError):
        device_ptr(torch.zeros(4).__dlpack__())            # a CPU DLPack capsule
    assert device_ptr(None) is None and device_ptr(1234) == 1234


def test_image_shard():
    from bayes_od_rc_b200.sharding import image_shard
    assert [image_shard(32, 8, r) for r in range(8)] == [(4 * r, 4) for r in range(8)]
    assert [image_shard(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert image_shard(2, 4, 3) == (2, 0)
    tot = sum(image_shard(129, 8, r)[1] for r in range(8))
    assert tot == 129


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["BOD_ROOT"])
import torch, torch.distributed as dist
from bayes_od_rc_b200.sharding import allgather_detections, image_shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
B, D, K = 3, 5, 4
first, count = image_shard(6, world, rank)
assert (first, count) == (3 * rank, 3)
blocks = dict(num_dets=torch.full((B,), rank + 1, dtype=torch.int32),
              means=torch.arange(B * D * 4, dtype=torch.float32).reshape(B, D, 4) + 1000 * rank,
              cat_param=torch.full((B, D, K), float(rank)))
g = allgather_detections(blocks, world)
assert g["num_dets"].tolist() == [1, 1, 1, 2, 2, 2]
assert g["means"].shape == (6, D, 4) and g["means"][3, 0, 0].item() == 1000.0 and g["means"][0, 0, 0].item() == 0.0
assert g["cat_param"][:3].eq(0).all() and g["cat_param"][3:].eq(1).all()
# timing reduction used by bench.py: max over ranks
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, BOD_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_strict_kernels_have_no_fma(tmp_path):
    """The translation units that implement the bit-exact arithmetic contract are
    compiled with -fmad=false: their PTX must not contain a binary32 fused
    multiply-add (IEEE division shows up as div.rn.f32; its internal Newton steps
    exist only in SASS and round correctly)."""
    from bayes_od_rc_b200 import build
    for unit in ("k2_posterior", "k3_softnms", "k4_fusion"):
        ptx = tmp_path / (unit + ".ptx")
        cmd = [build._nvcc()] + build.ARCH[:1] + ["arch=compute_100a,code=compute_100a"] + build.COMMON + \
            build.UNITS[unit + ".cu"] + ["-ptx", os.path.join(build.CSRC, unit + ".cu"), "-o", str(ptx)]
        subprocess.run(cmd, check=True, capture_output=True)
        txt = ptx.read_text()
        assert "fma.rn.f32" not in txt and "mad.f32" not in txt, f"{unit}: binary32 FMA in a strict kernel"
        assert "mul.f32" in txt or "mul.rn.f32" in txt
        # (the binary64 exp/log library routines use rcp.approx.ftz.f64 seeds internally; binary32 must be exact)
        assert not re.search(r"\.approx(\.ftz)?\.f32", txt), f"{unit}: approximate binary32 math in a strict kernel"


def test_k1_uses_bulk_copy_engine():
    from bayes_od_rc_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build.OBJDIR, "k1_moments.o")], capture_output=True,
                          text=True, check=True).stdout
    assert "UBLKCP" in sass, "k1 should move its tiles with cp.async.bulk (UBLKCP in SASS)"
    assert "SYNCS" in sass
