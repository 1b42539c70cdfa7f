"""CPU-only tests: the oracle against its committed goldens, the host logic of the library (uniform resolve, reference
BVH build, scene parser, band arithmetic) against the oracle, and that the C-ABI library loads and exports every symbol
include/rtb.h declares.  No compute call needs a GPU here."""
import ctypes as C
import hashlib
import json
import os
import sys
import re
import subprocess

import numpy as np
import pytest

from util import GOLDEN, REFERENCE_SCENES, abi, oracle_scene, params, scene_mod, synth, tiny_scene

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- the oracle is pinned by its goldens -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_oracle_reproduces_goldens(oracle, name):
    summary = json.load(open(os.path.join(GOLDEN, "summary.json")))[name]
    osc, holder = oracle_scene(oracle, synth.sample_scene(name))
    assert (osc.n_triangles, osc.n_nodes, osc.max_leaf) == (summary["n_triangles"], summary["n_nodes"], summary["max_leaf"])
    vn, mat, cen = osc.triangles()
    nodes, perm = osc.bvh()
    assert sha(vn) == summary["sha_triangles"] and sha(mat) == summary["sha_materials"] and sha(cen) == summary["sha_centers"]
    assert sha(nodes) == summary["sha_nodes"] and sha(perm) == summary["sha_perm"]
    g = np.load(os.path.join(GOLDEN, name + "_c1.npz"))
    r = osc.render(params(320, 240, 3), want_aux=True, want_rgbf=True)
    assert (r["prim"] == g["prim"]).all() and (r["mat"] == g["mat"]).all()
    assert (r["t"].view(np.uint32) == g["t"].view(np.uint32)).all()
    assert (r["rgba8"] == g["rgba8"]).all() and sha(r["rgbf"]) == summary["sha_rgbf"]
    c = r["counters"]
    for k in ("rays_primary", "rays_continuation", "rays_shadow", "primary_hits", "nodes_visited", "tris_tested", "max_stack"):
        assert getattr(c, k) == summary[k], k
    assert (osc.render(params(320, 240, 3, 4))["rgba8"] == g["rgba8_aa4"]).all()


def test_oracle_thread_count_does_not_change_results(oracle):
    osc, holder = oracle_scene(oracle, synth.sample_scene("test_scene_1"))
    a = osc.render(params(160, 120, 4, 4), threads=1)
    b = osc.render(params(160, 120, 4, 4), threads=0)
    assert (a["rgba8"] == b["rgba8"]).all() and a["counters"].rays == b["counters"].rays


def test_oracle_closest_hit_agrees_with_brute_force(oracle):
    """The BVH traversal of the oracle finds the same closest t as testing every triangle (float64-free property check)."""
    osc, holder = oracle_scene(oracle, synth.sample_scene("eval_scene"))
    p = params(64, 48, 1)
    r = osc.render(p, want_aux=True)
    for y in range(0, 48, 5):
        for x in range(0, 64, 7):
            o, d = osc.primary_ray(p, x, y)
            t, ids, n = osc.brute_closest(o, d)
            if n == 0:
                assert r["prim"][y, x] == -1
            else:
                assert np.float32(t).view(np.uint32) == r["t"][y, x].view(np.uint32) and r["prim"][y, x] in ids


@pytest.mark.parametrize("name", ["eval_scene", "test_scene_2"])  # leaves of 578 / 83 triangles (BVHBuilder.cs:142-145)
def test_oracle_leaf_accelerator_changes_nothing(oracle, name):
    """The checker-only accelerator for oversized leaves (oracle.cpp: LeafAccel) must return what the reference's linear leaf
    scan returns: frames, float accumulators, primary ids / t bits / materials and every counter, bit for bit."""
    packed = scene_mod.pack_scene(synth.sample_scene(name))
    prev = oracle.set_leaf_accel(0)
    try:
        plain = oracle.OracleScene.from_desc(packed.desc)
        oracle.set_leaf_accel(8)
        fast = oracle.OracleScene.from_desc(packed.desc)
    finally:
        oracle.set_leaf_accel(prev)
    assert plain.n_accelerated_leaves == 0 and fast.n_accelerated_leaves >= 5
    for kw in (dict(depth=6), dict(depth=4, aa=4, soft_shadows=1, light_size=5.0, glossy=1, roughness=0.05), dict(depth=3, is_orthographic=1)):
        p = params(200, 150, **kw)
        a, b = plain.render(p, want_aux=True, want_rgbf=True), fast.render(p, want_aux=True, want_rgbf=True)
        for key in ("rgba8", "prim", "mat"):
            assert (a[key] == b[key]).all(), (name, kw, key)
        assert (a["t"].view(np.uint32) == b["t"].view(np.uint32)).all() and (a["rgbf"].view(np.uint32) == b["rgbf"].view(np.uint32)).all()
        for f in ("rays_primary", "rays_continuation", "rays_shadow", "nodes_visited", "tris_tested", "closest_hits", "primary_hits"):
            assert getattr(a["counters"], f) == getattr(b["counters"], f), (name, kw, f)


def test_oracle_reaches_c3_at_full_size(oracle):
    """With the accelerator the degenerate reference BVH of the C3 sphere grid (two leaves of 98 310 triangles) renders 4K rows
    in a fraction of a second each; one row agrees with the plain linear scan (which needs ~30 s for it)."""
    packed = scene_mod.pack_scene(synth.sphere_grid_scene(16))
    osc = oracle.OracleScene.from_desc(packed.desc)
    assert osc.max_leaf == 98310 and osc.n_accelerated_leaves == 2
    p = params(3840, 2160, 16)
    fast = osc.render(p, rows=(1000, 1001, 1), want_aux=True)
    assert fast["counters"].seconds < 5.0
    prev = oracle.set_leaf_accel(0)
    try:
        plain_scene = oracle.OracleScene.from_desc(packed.desc)
    finally:
        oracle.set_leaf_accel(prev)
    pl = params(1280, 720, 2)  # a 720p row at depth 2 keeps the plain scan (98 310 triangle tests per node visit) within the CPU suite's budget
    a = plain_scene.render(pl, rows=(333, 334, 1), want_aux=True)
    b = osc.render(pl, rows=(333, 334, 1), want_aux=True)
    assert (a["rgba8"][333] == b["rgba8"][333]).all() and (a["prim"][333] == b["prim"][333]).all()
    assert (a["t"][333].view(np.uint32) == b["t"][333].view(np.uint32)).all()
    assert (a["prim"][333] >= 0).mean() > 0.3  # the row crosses the sphere grid


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SCENES), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_packed_scenes_match_reference_files(oracle, name):
    a = oracle.OracleScene.from_file(os.path.join(REFERENCE_SCENES, name + ".txt"))
    b, holder = oracle_scene(oracle, synth.sample_scene(name))
    for x, y in zip(a.triangles(), b.triangles()):
        assert x.tobytes() == y.tobytes()
    p = params(96, 72, 3)
    assert (a.frame(p) == b.frame(p)).all()


# ---- C ABI --------------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(pkg):
    lib = abi.load()
    header = open(os.path.join(ROOT, "include", "rtb.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void\*?|const [a-z_ ]+\*)\s+(rtb_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 37
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in rtb.h but not exported"
    assert declared == set(abi.SYMBOLS), declared ^ set(abi.SYMBOLS)
    assert lib.rtb_api_version() == 1


def test_struct_sizes_match_the_library(pkg):
    lib = abi.load()
    sizes = (C.c_int32 * 8)()
    lib.rtb_abi_sizes(sizes, 8)
    mine = [C.sizeof(t) for t in (abi.XformElem, abi.Material, abi.Triangle, abi.Mesh, abi.Prim, abi.SceneDesc, abi.RenderParams, abi.Stats)]
    assert list(sizes) == mine


def test_params_default_matches_reference_ui_defaults(pkg):
    lib = abi.load()
    p = abi.RenderParams()
    lib.rtb_params_default(C.byref(p))
    q = abi.default_params()
    assert bytes(p) == bytes(q)
    assert bytes(scene_mod.RenderSettings().to_params()) == bytes(q)


def test_no_gpu_means_an_error_not_a_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    ctx = C.c_void_p()
    assert lib.rtb_create(C.byref(ctx), None, 0) == abi.RTB_E_CUDA
    assert b"no CPU fallback" in lib.rtb_last_error(None)
    rt_mod = __import__("importlib").import_module("cosig-raytracing_b200.raytracer")
    with pytest.raises(rt_mod.RtbError):
        rt_mod.RayTracer()


# ---- host logic against the oracle -----------------------------------------------------------------------------------------
SETTINGS = [
    dict(), dict(has_fov=1, fov_deg=47.5), dict(has_bg=1, bg=(0.3, 0.5, 0.7)),
    dict(has_cam_pos=1, cam_pos=(3.0, -50.0, 20.0)), dict(has_cam_rot=1, cam_rot_euler_deg=(-35.0, 20.0, 170.0)),
    dict(has_cam_pos=1, cam_pos=(1.0, 2.0, 3.0), has_cam_rot=1, cam_rot_euler_deg=(10.0, 350.0, -45.0)),
    dict(is_orthographic=1, aa_samples=5),
]


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
@pytest.mark.parametrize("kw", SETTINGS)
def test_resolve_frame_bit_identical_to_oracle(pkg, oracle, name, kw):
    lib = abi.load()
    obj = synth.sample_scene(name)
    osc, holder = oracle_scene(oracle, obj)
    for has_res in (1, 0):
        p = params(**kw)
        p.has_resolution = has_res
        out = np.zeros(25, np.float32)
        wh = (C.c_int32 * 2)()
        assert lib.rtb_resolve_frame(holder.ptr(), C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
        assert out.view(np.uint32).tobytes() == osc.frame(p).view(np.uint32).tobytes()
        assert (wh[0], wh[1]) == osc.resolve(p)


def test_resolve_frame_defaults_without_image_camera_light(pkg, oracle):
    lib = abi.load()
    s = tiny_scene(3)
    s.Image, s.Camera, s.Lights = None, None, []
    osc, holder = oracle_scene(oracle, s)
    p = abi.default_params()
    out = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert lib.rtb_resolve_frame(holder.ptr(), C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    assert (wh[0], wh[1]) == (256, 256) == osc.resolve(p)
    assert out.tobytes() == osc.frame(p).tobytes()
    assert tuple(out[22:25]) == (np.float32(0.2),) * 3 and tuple(out[19:22]) == (0.0, 0.0, 0.0)


def test_resolve_frame_rejects_bad_parameters(pkg):
    lib = abi.load()
    holder = scene_mod.pack_scene(tiny_scene(1))
    for kw in (dict(depth=65), dict(depth=-1), dict(width=0), dict(aa=5000), dict(band_world=2, band_rank=2), dict(band_world=2, band_rows=6)):
        assert lib.rtb_resolve_frame(holder.ptr(), C.byref(params(**kw)), None, None) == abi.RTB_E_ARG, kw


def _raw12(vn, cen):
    raw = np.zeros((vn.shape[0], 12), np.float32)
    for k in range(3):
        raw[:, 4 * k:4 * k + 3] = vn[:, 3 * k:3 * k + 3]
        raw[:, 4 * k + 3] = cen[:, k]
    return raw


@pytest.mark.parametrize("scene_name", list(synth.SAMPLE_SCENES) + ["heightfield", "spheres", "tiny1", "tiny5"])
def test_reference_bvh_builder_matches_oracle(pkg, oracle, scene_name):
    lib = abi.load()
    obj = {"heightfield": lambda: synth.heightfield_scene(40, 30), "spheres": lambda: synth.sphere_grid_scene(3),
           "tiny1": lambda: tiny_scene(1), "tiny5": lambda: tiny_scene(5)}.get(scene_name, lambda: synth.sample_scene(scene_name))()
    osc, holder = oracle_scene(oracle, obj)
    vn, mat, cen = osc.triangles()
    raw = _raw12(vn, cen)
    n = raw.shape[0]
    nodes = np.zeros((2 * n + 1, 8), np.float32)
    perm = np.zeros(n, np.int32)
    nn = C.c_int64()
    assert lib.rtb_build_reference_bvh(raw.ctypes.data, n, nodes.ctypes.data, nodes.shape[0], C.byref(nn), perm.ctypes.data) == abi.RTB_OK
    onodes, operm = osc.bvh()
    assert nn.value == onodes.shape[0]
    assert nodes[:nn.value].view(np.uint32).tobytes() == onodes.view(np.uint32).tobytes()
    assert (perm == operm).all()


# ---- scene text parser --------------------------------------------------------------------------------------------------------
def _same_scene(a: scene_mod.ObjectData, b: scene_mod.ObjectData):
    pa, pb = scene_mod.pack_scene(a), scene_mod.pack_scene(b)
    assert pa.xoff.tobytes() == pb.xoff.tobytes() and bytes(pa.xel)[:20 * max(0, pa.xoff[-1])] == bytes(pb.xel)[:20 * max(0, pb.xoff[-1])]
    assert pa.tri.tobytes() == pb.tri.tobytes() and pa.mats.tobytes() == pb.mats.tobytes()
    assert pa.sph.tobytes() == pb.sph.tobytes() and pa.box.tobytes() == pb.box.tobytes()
    assert pa.lxf.tobytes() == pb.lxf.tobytes() and pa.lrgb.tobytes() == pb.lrgb.tobytes()
    for f in ("has_image", "image_w", "image_h", "has_camera", "cam_xform", "cam_distance", "cam_vfov_deg", "n_xforms", "n_lights",
              "n_materials", "n_meshes", "n_triangles", "n_spheres", "n_boxes"):
        assert getattr(pa.desc, f) == getattr(pb.desc, f), f
    assert tuple(pa.desc.bg) == tuple(pb.desc.bg)


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_native_parser_round_trips_sample_scenes(pkg, oracle, name):
    rt_mod = __import__("importlib").import_module("cosig-raytracing_b200.raytracer")
    obj = synth.sample_scene(name)
    text = synth.scene_to_text(obj).encode()
    parsed = rt_mod.SceneService.ParseScene(text)
    _same_scene(obj, parsed)
    # and the oracle's own restatement of SceneService reads the same text the same way
    a = oracle.OracleScene.from_text(text)
    b, holder = oracle_scene(oracle, parsed)
    for x, y in zip(a.triangles(), b.triangles()):
        assert x.tobytes() == y.tobytes()


def test_native_parser_round_trips_random_scenes(pkg, oracle):
    """The generator of tools/parity_fuzz.py (random transformations, materials, meshes, spheres, boxes; out-of-range indices; scenes
    without lights or materials) through text -> library parser -> same scene, and through the oracle's restatement of SceneService ->
    the same flattened triangles and the same rendered frame as the scene that never was text."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_fuzz as F
    rt_mod = __import__("importlib").import_module("cosig-raytracing_b200.raytracer")
    for seed in range(40):
        obj = F.random_scene(seed)
        text = synth.scene_to_text(obj).encode()
        parsed = rt_mod.SceneService.ParseScene(text)
        _same_scene(obj, parsed)
        a = oracle.OracleScene.from_text(text)
        b, holder = oracle_scene(oracle, obj)
        assert a.triangles()[0].shape == b.triangles()[0].shape
        for x, y in zip(a.triangles(), b.triangles()):
            assert x.tobytes() == y.tobytes(), seed
        if seed % 8 == 0:
            p, _ = F.random_settings(seed)
            assert (a.render(p)["rgba8"] == b.render(p)["rgba8"]).all(), seed


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SCENES), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_native_parser_reads_reference_files(pkg, name):
    rt_mod = __import__("importlib").import_module("cosig-raytracing_b200.raytracer")
    _same_scene(rt_mod.SceneService.LoadScene(os.path.join(REFERENCE_SCENES, name + ".txt")), synth.sample_scene(name))


def test_parser_tolerances_and_errors(pkg):
    rt_mod = __import__("importlib").import_module("cosig-raytracing_b200.raytracer")
    text = (b"// a comment line\r\n\r\nIMAGE\r\n\r\n{\r\n\t 32  24 // res\r\n 0.1\t0.2 0.3\r\n}\r\n"
            b"transformation\n{\n T 1 2 3\n Bogus 1 2\n\n Rz 1e1\n}\n"
            b"Unknown\n{\n 1 2 3\n}\n"
            b"Camera\n{\n0\n30\n45.5\n}\nLight\n{\n0\n1 1 1\n}\nMaterial\n{\n1 0 0\n0.1 0.2 0.3 0.4 1.5\n}\n"
            b"Triangles\n{\n0\n\n0\n0 0 0\n1 0 0\n0 1 0\n}\nSphere\n{\n0\n0\n}\nbox\n{\n0\n0\n}")
    s = rt_mod.SceneService.ParseScene(text)
    assert (s.Image.horizontal, s.Image.vertical) == (32, 24) and s.Image.background == pytest.approx((0.1, 0.2, 0.3))
    assert len(s.Transformations) == 1 and [e.Type for e in s.Transformations[0].Elements] == [abi.RTB_XF_T, abi.RTB_XF_RZ]
    assert s.Transformations[0].Elements[1].AngleDeg == 10.0
    assert s.Camera.verticalFovDeg == 45.5 and len(s.Lights) == 1 and len(s.Materials) == 1 and s.Materials[0].ior == 1.5
    assert s.TriangleMeshes[0].materials.tolist() == [0] and len(s.Spheres) == 1 and len(s.Boxes) == 1
    with pytest.raises(rt_mod.RtbError) as e:
        rt_mod.SceneService.ParseScene(b"Camera\n{\n0\nabc\n30\n}\n")
    assert e.value.code == abi.RTB_E_PARSE
    with pytest.raises(rt_mod.RtbError):
        rt_mod.SceneService.ParseScene(b"Triangles\n{\n0\n1\n0 0 0\n1 0 0\n")  # truncated triangle
    empty = rt_mod.SceneService.LoadScene("/nonexistent/scene.txt")  # missing file -> empty ObjectData (SceneService.cs:28-33)
    assert empty.Image is None and not empty.TriangleMeshes
    assert not rt_mod.SceneService.ParseScene(b"").Transformations


def test_exact_closest_mode_is_the_brute_force_closest_hit(oracle):
    """Checker-only mode of the oracle (orc_set_exact_closest) that the LBVH flavour is held against: on random scenes its primary hits
    must carry the t bits of a scan over ALL triangles, also on the pixels where the reference's own traversal (FP32 cull on exact boxes)
    keeps the farther of two nearly coincident surfaces — seed 1199 of tools/parity_fuzz.py has 15 of those."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_fuzz as F
    checked = differing = 0
    for seed in (3, 7, 19, 1199, 1598):
        obj = F.random_scene(seed)
        p, _ = F.random_settings(seed)
        packed = scene_mod.pack_scene(obj)
        osc = oracle.OracleScene.from_desc(packed.desc)
        ref = osc.render(p, want_aux=True)
        osc.set_exact_closest(True)
        rx = osc.render(p, want_aux=True)
        osc.set_exact_closest(False)
        again = osc.render(p, want_aux=True)
        assert (again["rgba8"] == ref["rgba8"]).all()  # the switch leaves the reference's traversal as it was
        h, w = rx["prim"].shape
        ys, xs = np.nonzero((rx["t"].view(np.uint32) != ref["t"].view(np.uint32)) | (rx["prim"] != ref["prim"]))
        differing += len(ys)
        rng = np.random.RandomState(seed)
        for y, x in list(zip(ys, xs)) + [(rng.randint(h), rng.randint(w)) for _ in range(120)]:
            o, d = osc.primary_ray(p, int(x), int(y))
            tb, ids, n = osc.brute_closest(o, d, cap=64)
            checked += 1
            if n == 0:
                assert rx["prim"][y, x] < 0
            else:
                assert rx["prim"][y, x] >= 0 and np.float32(tb).view(np.uint32) == rx["t"].view(np.uint32)[y, x], (seed, x, y)
    assert differing >= 15 and checked > 600


# ---- band arithmetic and the world-size-2 gather (gloo) ------------------------------------------------------------------------
def test_band_partition_covers_every_row_once():
    bands = __import__("importlib").import_module("cosig-raytracing_b200.bands")
    for h in (1, 31, 32, 33, 150, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            for band_rows in (4, 8, 16, 32):
                seen = np.zeros(h, np.int32)
                for rank in range(world):
                    rows = bands.owned_rows(h, rank, world, band_rows)
                    assert len(rows) == bands.local_row_count(h, rank, world, band_rows)
                    seen[rows] += 1
                assert (seen == 1).all()


def _gloo_worker(rank, world, port, h, w, out_q, band_rows=32):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bands = importlib.import_module("cosig-raytracing_b200.bands")
    full = (np.arange(h * w * 4, dtype=np.int64) % 251).astype(np.uint8).reshape(h, w, 4)  # the frame every rank would render
    mine = torch.from_numpy(full[bands.owned_rows(h, rank, world, band_rows)].copy())
    frame = bands.gather_bands(mine, h, w, rank, world, band_rows, dst=0)
    if rank == 0:
        out_q.put(bool((frame.numpy() == full).all()))
    dist.destroy_process_group()


@pytest.mark.parametrize("h,band_rows", [(64, 32), (150, 32), (150, 8)])  # bench.py shards by 8-row bands
def test_gather_bands_world_size_2_gloo(h, band_rows):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + h + band_rows) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, h, 48, q, band_rows)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_host_ring_copy_plans_tile_the_frame_once():
    """rtb_group_create_host: every rank copies `full` strided pieces plus at most one short tail.  Over all ranks the plans must write
    every byte of the frame exactly once and move exactly the rank's rows (ragged heights, more ranks than bands, world = 1)."""
    bands = __import__("importlib").import_module("cosig-raytracing_b200.bands")
    for h in (1, 7, 8, 9, 31, 33, 150, 236, 2160):
        for world in (1, 2, 3, 4, 8):
            for band_rows in (4, 8, 16, 32):
                w = 5
                src = (np.arange(h * w * 4, dtype=np.int64) % 251 + 1).astype(np.uint8)
                dst = np.zeros_like(src)
                hits = np.zeros(src.size, np.int32)
                for rank in range(world):
                    plan = bands.host_ring_copy_plan(h, w, rank, world, band_rows)
                    moved = bands.apply_copy_plan(plan, src, dst)
                    assert moved == bands.local_row_count(h, rank, world, band_rows) * w * 4, (h, world, band_rows, rank)
                    mark = np.zeros_like(src)
                    bands.apply_copy_plan(plan, np.ones_like(src), mark)
                    hits += mark
                    rows = bands.owned_rows(h, rank, world, band_rows)
                    assert (mark.reshape(h, w * 4)[rows] == 1).all() and mark.sum() == len(rows) * w * 4
                assert (hits == 1).all() and (dst == src).all(), (h, world, band_rows)


def _shm_ring_worker(rank, world, name, h, w, band_rows, frames, n_buf):
    """One rank of the host-ring protocol without a GPU: waits for rank 0's begin, writes its rows of frame k into slot k % n_buf,
    posts done[rank][slot] = k + 1 (the words and their meaning are api.cu's HostRingHeader)."""
    import importlib
    import time
    from multiprocessing import shared_memory
    bands = importlib.import_module("cosig-raytracing_b200.bands")
    shm = shared_memory.SharedMemory(name=name)
    try:
        words = np.ndarray((1024,), dtype=np.uint32, buffer=shm.buf)   # [0] = begun0, [16 + rank * 8 + slot] = done
        frame_bytes = h * w * 4
        plan = bands.host_ring_copy_plan(h, w, rank, world, band_rows)
        for k in range(frames):
            t0 = time.time()
            while int(words[0]) < k + 1:
                assert time.time() - t0 < 60
                time.sleep(0.0005)
            j = k % n_buf
            src = ((np.arange(frame_bytes, dtype=np.int64) + 7 * k) % 251).astype(np.uint8)   # frame k as every rank would render it
            dst = np.ndarray((frame_bytes,), dtype=np.uint8, buffer=shm.buf, offset=4096 + j * frame_bytes)
            bands.apply_copy_plan(plan, src, dst)
            words[16 + rank * 8 + j] = k + 1
        del words, dst
    finally:
        shm.close()


@pytest.mark.parametrize("world,h,band_rows", [(2, 150, 8), (3, 236, 16)])
def test_host_ring_protocol_over_shared_memory(world, h, band_rows):
    """The hand-shake of the host ring run by `world` CPU processes over POSIX shared memory (no CUDA): more frames than slots, rank 0
    begins frame k only after it has read frame k - n_buf, the others may write a slot only once rank 0 has begun that frame.  Every
    frame rank 0 reads must be complete and be the right one."""
    import multiprocessing as mp
    import time
    from multiprocessing import shared_memory
    bands = __import__("importlib").import_module("cosig-raytracing_b200.bands")
    w, frames, n_buf = 48, 9, 2
    frame_bytes = h * w * 4
    shm = shared_memory.SharedMemory(create=True, size=4096 + n_buf * frame_bytes)
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_shm_ring_worker, args=(r, world, shm.name, h, w, band_rows, frames, n_buf)) for r in range(1, world)]
    try:
        words = np.ndarray((1024,), dtype=np.uint32, buffer=shm.buf)
        words[:] = 0
        for p in procs:
            p.start()
        plan = bands.host_ring_copy_plan(h, w, 0, world, band_rows)

        def read(k):
            j = k % n_buf
            t0 = time.time()
            while any(int(words[16 + r * 8 + j]) < k + 1 for r in range(world)):
                assert time.time() - t0 < 60, "a rank did not arrive"
                time.sleep(0.0005)
            got = np.ndarray((frame_bytes,), dtype=np.uint8, buffer=shm.buf, offset=4096 + j * frame_bytes)
            want = ((np.arange(frame_bytes, dtype=np.int64) + 7 * k) % 251).astype(np.uint8)
            assert (got == want).all(), f"frame {k}"

        for k in range(frames):
            if k >= n_buf:
                read(k - n_buf)
            words[0] = k + 1                                    # rank 0 begins frame k: slot k % n_buf may be overwritten now
            j = k % n_buf
            src = ((np.arange(frame_bytes, dtype=np.int64) + 7 * k) % 251).astype(np.uint8)
            dst = np.ndarray((frame_bytes,), dtype=np.uint8, buffer=shm.buf, offset=4096 + j * frame_bytes)
            bands.apply_copy_plan(plan, src, dst)
            words[16 + 0 * 8 + j] = k + 1
        for k in range(max(0, frames - n_buf), frames):
            read(k)
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        del words, dst
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
        shm.close()
        shm.unlink()


# ---- the C++ host mirror (include/rtb_raytracer.hpp) --------------------------------------------------------------------------
def build_cpp_test():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def test_cpp_host_mirror_builds_and_runs_host_checks(pkg, tmp_path):
    import subprocess
    exe = build_cpp_test()
    scene_file = tmp_path / "scene.txt"
    scene_file.write_bytes(synth.scene_to_text(synth.sample_scene("test_scene_2")).encode())
    r = subprocess.run([exe, "host", str(scene_file)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "host OK 8 transformations" in r.stdout


# ---- bench.py contract: one JSON line on stdout ------------------------------------------------------------------------------------
def build_binding_sequence():
    """tests/c/binding_sequence.c: the call sequence of csharp/RayTracerNative.cs in plain C (gcc, C11) against include/rtb.h."""
    exe = os.path.join(ROOT, "tests", "c", "binding_sequence")
    src = os.path.join(ROOT, "tests", "c", "binding_sequence.c")
    lib_dir = os.path.join(ROOT, "cosig-raytracing_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", src, "-I", os.path.join(ROOT, "include"), "-L", lib_dir, "-lrtb200",
                           f"-Wl,-rpath,{lib_dir}", "-o", exe])
    return exe


def test_c_binding_sequence_host_part(pkg):
    """The header is usable from C (not only C++), and the host-only calls the C# static constructor / marshalling make behave."""
    exe = build_binding_sequence()
    r = subprocess.run([exe, "host"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr


def test_csharp_binding_declares_what_it_calls():
    """csharp/RayTracerNative.cs cannot be compiled here; at least every entry point it imports must exist in rtb.h with the same
    number of parameters, and it must actually upload scenes and define the reference's public surface."""
    src = open(os.path.join(ROOT, "csharp", "RayTracerNative.cs")).read()
    header = open(os.path.join(ROOT, "include", "rtb.h")).read()
    imports = re.findall(r"\[DllImport\(Lib\)\]\s+public static extern (?:unsafe )?[\w\*]+ (rtb_\w+)\(([^)]*)\)", src)
    assert len(imports) >= 22
    for name, args in imports:
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\);", header, re.S)
        assert m, f"{name} is imported by the C# binding but not declared in rtb.h"
        n_cs = 0 if not args.strip() else len(args.split(","))
        n_c = 0 if m.group(1).strip() in ("", "void") else len(m.group(1).split(","))
        assert n_cs == n_c, f"{name}: {n_cs} parameters in C#, {n_c} in rtb.h"
    body = src[src.index("unsafe void EnsureScene"):]
    assert "RtbNative.rtb_upload_scene(ctx, &d, PrimitiveMode, BvhMode)" in body and "omitted" not in src
    for member in ("SetComputeShader", "InvalidateBVHCache", "ReleaseBuffers", "ClearRenderTarget", "RenderTexture RenderToTexture(ObjectData scene, RenderSettings settings)",
                   "Task<Texture2D> RenderAsync(ObjectData scene, RenderSettings settings, IProgress<float> progress, CancellationToken token)", "SaveTexture",
                   "RenderBegin", "RenderEnd", "GetStats"):
        assert member in src, member


def test_bench_reference_arm_prints_exactly_one_json_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores; needs no GPU) must put ONE JSON line on stdout with the
    driver's keys; everything else (library banners, warnings) goes to stderr."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
