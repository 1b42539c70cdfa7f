"""GPU parity tests: the CUDA path (through the C ABI of librtb200.so) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): integer / index work bit-exact — flattened triangle arrays, BVH nodes, primary-hit
prim ids, t bits and material ids, ray counts; final RGB within 1/255 per channel on >= 99.9 % of pixels (in parity
mode the frames are in fact required to be identical, because the arithmetic spec is shared).
"""
import importlib
import os

import numpy as np
import pytest

from util import abi, assert_rgb_parity, oracle_scene, params, rgb_agreement, scene_mod, synth, tiny_scene

pytestmark = pytest.mark.gpu

rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")


@pytest.fixture(scope="module")
def tracers(pkg):
    made = {}

    def get(mode):
        if mode not in made:
            made[mode] = rt_mod.RayTracer(bvh_mode=mode)
        return made[mode]

    yield get
    for rt in made.values():
        rt.close()


@pytest.fixture(scope="module")
def samples(oracle):
    out = {}
    for name in synth.SAMPLE_SCENES:
        obj = synth.sample_scene(name)
        osc, holder = oracle_scene(oracle, obj)
        out[name] = (obj, osc, holder)
    return out


# ---- K1: scene upload -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_flattened_triangles_bit_identical(tracers, samples, name):
    obj, osc, _ = samples[name]
    rt = tracers(abi.RTB_BVH_REFERENCE)
    rt.RenderToTexture(obj, params(16, 16, 1))
    vn, mat = rt.triangles()
    ovn, omat, _ = osc.triangles()
    assert vn.shape == ovn.shape
    assert vn.view(np.uint32).tobytes() == ovn.view(np.uint32).tobytes()
    assert (mat == omat).all()


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_reference_bvh_identical(tracers, samples, name):
    obj, osc, _ = samples[name]
    rt = tracers(abi.RTB_BVH_REFERENCE)
    rt.RenderToTexture(obj, params(16, 16, 1))
    nodes, perm = rt.bvh()
    onodes, operm = osc.bvh()
    assert nodes.shape == onodes.shape
    assert nodes.view(np.uint32).tobytes() == onodes.view(np.uint32).tobytes()
    assert (perm == operm).all()


# ---- K3/K4: primary rays ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
@pytest.mark.parametrize("res", [(320, 240), (1920, 1080)])
def test_primary_hits_bit_exact_reference_mode(tracers, samples, name, res):
    obj, osc, _ = samples[name]
    p = params(res[0], res[1], 3)
    prim, t, mat = tracers(abi.RTB_BVH_REFERENCE).primary_hits(obj, p)
    ref = osc.render(p, want_aux=True)
    assert (prim == ref["prim"]).all()
    assert (t.view(np.uint32) == ref["t"].view(np.uint32)).all()
    assert (mat == ref["mat"]).all()


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_primary_hits_match_golden(tracers, samples, name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + "_c1.npz"))
    prim, t, mat = tracers(abi.RTB_BVH_REFERENCE).primary_hits(samples[name][0], params(320, 240, 3))
    assert (prim == g["prim"]).all() and (mat == g["mat"]).all()
    assert (t.view(np.uint32) == g["t"].view(np.uint32)).all()


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_primary_hits_lbvh_mode(tracers, samples, oracle, name):
    """LBVH mode: t bits and material equal everywhere the closest hit is unique; prim ids may differ only where several
    triangles attain exactly the closest t (the winner then depends on traversal order)."""
    obj, osc, _ = samples[name]
    p = params(640, 480, 3)
    prim, t, mat = tracers(abi.RTB_BVH_LBVH).primary_hits(obj, p)
    ref = osc.render(p, want_aux=True)
    same_t = t.view(np.uint32) == ref["t"].view(np.uint32)
    assert same_t.mean() >= 0.9999, f"t mismatch on {(~same_t).sum()} pixels"
    differ = np.argwhere((prim != ref["prim"]) & same_t)
    for y, x in differ[:200]:
        o, d = osc.primary_ray(p, int(x), int(y))
        tb, ids, n = osc.brute_closest(o, d, cap=64)
        assert n >= 2 and prim[y, x] in ids, f"pixel ({x},{y}): id {prim[y, x]} vs {ref['prim'][y, x]} without a tie"
    assert len(differ) <= 0.002 * prim.size


# ---- full frames --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
@pytest.mark.parametrize("depth,aa", [(3, 1), (6, 1), (2, 4), (1, 1)])
def test_frame_parity_reference_mode(tracers, samples, name, depth, aa):
    obj, osc, _ = samples[name]
    p = params(320, 240, depth, aa)
    rt = tracers(abi.RTB_BVH_REFERENCE)
    tex = rt.RenderAsync(obj, p)
    ref = osc.render(p)
    within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"{name} d{depth} aa{aa}")
    assert same == 1.0, f"parity mode shares the arithmetic spec: frames must be identical (got {same}, worst {worst})"
    s, c = rt.stats(), ref["counters"]
    assert (s.rays_primary, s.rays_continuation, s.rays_shadow) == (c.rays_primary, c.rays_continuation, c.rays_shadow)
    assert s.paths_hit_primary == c.primary_hits
    assert s.reserved[0] == 0  # no traversal-stack overflow


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_frame_matches_golden(tracers, samples, name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + "_c1.npz"))
    rt = tracers(abi.RTB_BVH_REFERENCE)
    assert (rt.RenderAsync(samples[name][0], params(320, 240, 3)).pixels == g["rgba8"]).all()
    assert (rt.RenderAsync(samples[name][0], params(320, 240, 3, 4)).pixels == g["rgba8_aa4"]).all()


def test_c2_full_size(tracers, samples):
    """BASELINE config C2: sample scene, 1920x1080, depth 6."""
    obj, osc, _ = samples["test_scene_1"]
    p = params(1920, 1080, 6)
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        tex = tracers(mode).RenderAsync(obj, p)
        ref = osc.render(p)
        within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"C2 mode {mode}")
        if mode == abi.RTB_BVH_REFERENCE:
            assert same == 1.0


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_frame_parity_lbvh_mode(tracers, samples, name):
    obj, osc, _ = samples[name]
    p = params(640, 480, 6)
    tex = tracers(abi.RTB_BVH_LBVH).RenderAsync(obj, p)
    assert_rgb_parity(tex.pixels, osc.render(p)["rgba8"], f"{name} lbvh")


@pytest.mark.parametrize("kw", [
    dict(is_orthographic=1),
    dict(soft_shadows=1, light_size=5.0, aa_samples=4),
    dict(glossy=1, roughness=0.05, aa_samples=4),
    dict(motion_blur=1, shutter_speed=1.0, aa_samples=4),
    dict(enable_ambient=0), dict(enable_diffuse=0), dict(enable_specular=0), dict(enable_refraction=0),
    dict(light_intensity=1.7), dict(has_bg=1, bg=(0.9, 0.1, 0.4)), dict(has_fov=1, fov_deg=55.0),
    dict(has_cam_pos=1, cam_pos=(5.0, -60.0, 30.0), has_cam_rot=1, cam_rot_euler_deg=(-60.0, 10.0, 5.0)),
    dict(has_cam_rot=1, cam_rot_euler_deg=(0.0, 0.0, 30.0)),
    dict(debug_mode=1), dict(debug_mode=2), dict(debug_mode=3), dict(debug_mode=2, is_orthographic=1),
    dict(aa_samples=3), dict(aa_samples=8), dict(aa_samples=16),
])
def test_render_settings_parity(tracers, samples, kw):
    obj, osc, _ = samples["test_scene_2"]
    p = params(200, 152, 4, **kw)
    tex = tracers(abi.RTB_BVH_REFERENCE).RenderAsync(obj, p)
    within, same, worst = assert_rgb_parity(tex.pixels, osc.render(p)["rgba8"], str(kw))
    assert same == 1.0, f"{kw}: identical {same}, worst {worst}"


# ---- edge cases -----------------------------------------------------------------------------------------------------------
def test_empty_scene_renders_background(tracers, oracle):
    s = scene_mod.ObjectData()
    synth._sample_camera_and_light(s)
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        tex = tracers(mode).RenderAsync(s, params(40, 24, 3))
        assert (tex.pixels[..., :3] == 51).all() and (tex.pixels[..., 3] == 255).all()  # 0.2 * 255 = 51
        prim, t, mat = tracers(mode).primary_hits(s, params(40, 24, 3))
        assert (prim == -1).all()


def test_zero_depth_is_black(tracers, samples):
    obj, osc, _ = samples["test_scene_1"]
    tex = tracers(abi.RTB_BVH_REFERENCE).RenderAsync(obj, params(64, 40, 0))
    assert (tex.pixels[..., :3] == 0).all() and (tex.pixels[..., 3] == 255).all()
    assert (osc.render(params(64, 40, 0))["rgba8"] == tex.pixels).all()


@pytest.mark.parametrize("n", [1, 2, 4, 5, 33])
def test_tiny_scenes_both_modes(tracers, oracle, n):
    s = tiny_scene(n)
    osc, holder = oracle_scene(oracle, s)
    p = params(96, 64, 3)
    ref = osc.render(p, want_aux=True)
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        rt = tracers(mode)
        tex = rt.RenderAsync(s, p)
        prim, t, mat = rt.primary_hits(s, p)
        assert_rgb_parity(tex.pixels, ref["rgba8"], f"tiny {n} mode {mode}")
        assert (t.view(np.uint32) == ref["t"].view(np.uint32)).mean() >= 0.999
        if mode == abi.RTB_BVH_REFERENCE:
            assert (prim == ref["prim"]).all() and (tex.pixels == ref["rgba8"]).all()


def test_material_index_out_of_range_uses_defaults(tracers, oracle):
    s = tiny_scene(6)
    s.TriangleMeshes[0].materials[:] = [0, -1, 7, 0, 99, -5]
    osc, holder = oracle_scene(oracle, s)
    p = params(96, 64, 2)
    tex = tracers(abi.RTB_BVH_REFERENCE).RenderAsync(s, p)
    assert (tex.pixels == osc.render(p)["rgba8"]).all()


def test_no_materials_no_lights_no_camera(tracers, oracle):
    s = tiny_scene(8)
    s.Materials, s.Lights, s.Camera, s.Image = [], [], None, None
    osc, holder = oracle_scene(oracle, s)
    p = params(64, 64, 2, has_cam_pos=1, cam_pos=(0.0, 0.0, 40.0))
    tex = tracers(abi.RTB_BVH_REFERENCE).RenderAsync(s, p)
    assert (tex.pixels == osc.render(p)["rgba8"]).all()
    p2 = abi.default_params()  # no resolution anywhere -> 256 x 256 (RayTracer.cs:221-222)
    tex2 = tracers(abi.RTB_BVH_REFERENCE).RenderAsync(s, p2)
    assert tex2.pixels.shape == (256, 256, 4)


def test_errors(tracers, samples, abi):
    import ctypes as C
    lib = abi.load()
    ctx = C.c_void_p()
    assert lib.rtb_create(C.byref(ctx), None, 0) == abi.RTB_OK
    p = params(32, 32, 2)
    buf = np.zeros((32, 32, 4), np.uint8)
    assert lib.rtb_render(ctx, C.byref(p), buf.ctypes.data, buf.nbytes, None, None) == abi.RTB_E_NOSCENE  # render before upload
    packed = scene_mod.pack_scene(samples["test_scene_1"][0])
    assert lib.rtb_upload_scene(ctx, packed.ptr(), abi.RTB_PRIM_TESSELLATED, abi.RTB_BVH_REFERENCE) == abi.RTB_OK
    assert lib.rtb_render(ctx, C.byref(p), buf.ctypes.data, 100, None, None) == abi.RTB_E_SIZE
    assert b"too small" in lib.rtb_last_error(ctx)
    bad = params(32, 32, 99)
    assert lib.rtb_render(ctx, C.byref(bad), buf.ctypes.data, buf.nbytes, None, None) == abi.RTB_E_ARG
    assert lib.rtb_render(ctx, C.byref(p), buf.ctypes.data, buf.nbytes, None, None) == abi.RTB_OK
    assert lib.rtb_invalidate(ctx) == abi.RTB_OK
    assert lib.rtb_render(ctx, C.byref(p), buf.ctypes.data, buf.nbytes, None, None) == abi.RTB_E_NOSCENE
    flag = C.c_int32(1)  # cancellation observed before the first chunk
    assert lib.rtb_upload_scene(ctx, packed.ptr(), abi.RTB_PRIM_TESSELLATED, abi.RTB_BVH_REFERENCE) == abi.RTB_OK
    lib.rtb_set_cancel_flag(ctx, C.addressof(flag))
    assert lib.rtb_render(ctx, C.byref(p), buf.ctypes.data, buf.nbytes, None, None) == abi.RTB_E_CANCELLED
    lib.rtb_set_cancel_flag(ctx, None)
    assert lib.rtb_upload_scene(ctx, packed.ptr(), 7, abi.RTB_BVH_REFERENCE) == abi.RTB_E_ARG  # unknown primitive mode
    lib.rtb_destroy(ctx)


def test_bvh_cache_semantics(tracers, samples):
    """RayTracer.cs:118-123: the scene is uploaded once per scene object; InvalidateBVHCache forces a rebuild."""
    obj = samples["test_scene_1"][0]
    rt = rt_mod.RayTracer()
    p = params(64, 48, 2)
    a = rt.RenderAsync(obj, p).pixels
    up0 = rt.stats().ms_upload
    b = rt.RenderAsync(obj, p).pixels
    assert rt.stats().ms_upload == up0 and (a == b).all()
    rt.InvalidateBVHCache()
    c = rt.RenderAsync(obj, p).pixels
    assert (a == c).all()
    rt.ReleaseBuffers()
    assert (rt.RenderAsync(obj, p).pixels == a).all()
    rt.close()


# ---- tile sharding and chunking (single GPU stands in for the ranks) ------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3, 8])
def test_band_sharding_reassembles_the_frame(tracers, samples, world):
    import torch
    obj, osc, _ = samples["test_scene_1"]
    w, h = 200, 150  # 150 rows: the last band is ragged
    rt = tracers(abi.RTB_BVH_REFERENCE)
    full = rt.RenderAsync(obj, params(w, h, 3)).pixels
    frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for rank in range(world):
        rt.RenderToTexture(obj, params(w, h, 3, band_rank=rank, band_world=world, band_rows=16), frame.data_ptr(), frame.numel())
    assert (frame.cpu().numpy() == full).all()
    # compact layout: each rank's bands packed, then scattered back on the host
    out = np.zeros_like(full)
    for rank in range(world):
        rows = [r for r in range(h) if (r // 16) % world == rank]
        buf = torch.zeros((max(1, len(rows)), w, 4), dtype=torch.uint8, device="cuda")
        rt.RenderToTexture(obj, params(w, h, 3, band_rank=rank, band_world=world, band_rows=16, out_layout=abi.RTB_OUT_COMPACT), buf.data_ptr(), buf.numel())
        if rows:
            out[rows] = buf.cpu().numpy()[:len(rows)]
    assert (out == full).all()


def test_chunked_frame_equals_single_chunk(samples, monkeypatch):
    obj = samples["test_scene_2"][0]
    p = params(320, 240, 4, 4)
    a = rt_mod.RayTracer()
    ref = a.RenderAsync(obj, p).pixels
    assert a.stats().chunks == 1
    a.close()
    monkeypatch.setenv("RTB_CHUNK_SLOTS", "20000")
    b = rt_mod.RayTracer()
    got = b.RenderAsync(obj, p).pixels
    assert b.stats().chunks > 4
    b.close()
    assert (got == ref).all()


# ---- larger synthetic scenes ------------------------------------------------------------------------------------------------
def test_heightfield_small_against_oracle(tracers, oracle):
    s = synth.heightfield_scene(100, 50)  # 10 000 triangles, same generator as C4
    osc, holder = oracle_scene(oracle, s)
    p = params(480, 270, 6)
    ref = osc.render(p, want_aux=True)
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        rt = tracers(mode)
        tex = rt.RenderAsync(s, p)
        prim, t, mat = rt.primary_hits(s, p)
        assert_rgb_parity(tex.pixels, ref["rgba8"], f"heightfield mode {mode}")
        assert (t.view(np.uint32) == ref["t"].view(np.uint32)).mean() >= 0.9999
        assert (mat == ref["mat"]).mean() >= 0.9999
        if mode == abi.RTB_BVH_REFERENCE:
            assert (prim == ref["prim"]).all() and (tex.pixels == ref["rgba8"]).all()
        assert rt.stats().reserved[0] == 0


def test_sphere_grid_small_against_oracle(tracers, oracle):
    s = synth.sphere_grid_scene(4)  # 16 spheres + floor: 12 300 triangles, glass + mirror, depth 16 like C3
    osc, holder = oracle_scene(oracle, s)
    p = params(480, 270, 16)
    ref = osc.render(p)
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        tex = tracers(mode).RenderAsync(s, p)
        within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"sphere grid mode {mode}")
        if mode == abi.RTB_BVH_REFERENCE:
            assert same == 1.0


def test_c4_full_size_properties(tracers, oracle):
    """BASELINE config C4 at full size (1 000 000 triangles, 3840x2160, depth 6): the oracle checks every 16th row;
    the two GPU BVH modes must agree with each other on the whole frame (size-independent property: the image does not
    depend on the acceleration structure)."""
    s = synth.heightfield_scene()
    p = params(3840, 2160, 6)
    lb = tracers(abi.RTB_BVH_LBVH)
    tex_l = lb.RenderAsync(s, p).pixels
    st = lb.stats()
    assert st.n_triangles == 1_000_000 and st.reserved[0] == 0
    rf = tracers(abi.RTB_BVH_REFERENCE)
    tex_r = rf.RenderAsync(s, p).pixels
    within, same, worst = rgb_agreement(tex_l, tex_r)
    assert within >= 0.999, (within, same, worst)
    osc, holder = oracle_scene(oracle, s)
    ref = osc.render(p, rows=(0, -1, 16))
    rows = np.arange(0, 2160, 16)
    assert (tex_r[rows] == ref["rgba8"][rows]).all()
    assert_rgb_parity(tex_l[rows], ref["rgba8"][rows], "C4 LBVH vs oracle rows")


# ---- the C++ host mirror drives the same path -----------------------------------------------------------------------------------
def test_cpp_host_mirror_render_matches_oracle(samples, tmp_path):
    import subprocess
    from test_host_cpu import build_cpp_test
    exe = build_cpp_test()
    obj, osc, _ = samples["test_scene_1"]
    scene_file, out = tmp_path / "scene.txt", tmp_path / "out.rgba"
    scene_file.write_bytes(synth.scene_to_text(obj).encode())
    r = subprocess.run([exe, "render", str(scene_file), str(out), "160", "120", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.uint8).reshape(120, 160, 4)
    assert (got == osc.render(params(160, 120, 4))["rgba8"]).all()


def test_c_binding_sequence_on_the_gpu(samples, tmp_path):
    """tests/c/binding_sequence.c walks the call sequence of csharp/RayTracerNative.cs (upload, RenderAsync with the cancel flag,
    RenderToTexture's begin / end, tickets in flight, invalidate / re-upload, clear target, a file scene, the rotation GIF) from
    plain C; the frame it renders from the scene file must be the oracle's."""
    import subprocess
    from test_host_cpu import build_binding_sequence
    exe = build_binding_sequence()
    obj, osc, _ = samples["test_scene_1"]
    scene_file, out = tmp_path / "scene.txt", tmp_path / "out.rgba"
    scene_file.write_bytes(synth.scene_to_text(obj).encode())
    r = subprocess.run([exe, "gpu", str(scene_file), str(out), "200", "150"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.uint8).reshape(150, 200, 4)
    assert (got == osc.render(params(200, 150, 3))["rgba8"]).all()
    assert os.path.getsize(str(out) + ".gif") > 1000


def test_multi_device_context_matches_single(samples):
    """One process driving two GPUs: bands over the devices, peers store into device 0's frame (needs >= 2 GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    obj = samples["test_scene_1"][0]
    p = params(640, 360, 4)
    one = rt_mod.RayTracer(devices=[0])
    ref = one.RenderAsync(obj, p).pixels
    one.close()
    two = rt_mod.RayTracer(devices=[0, 1])
    got = two.RenderAsync(obj, p).pixels
    st = two.stats()
    assert st.n_devices == 2 and (got == ref).all()
    # pipelined frames on the two-device context: more tickets in flight than device frame buffers
    settings = [params(640, 360, 4, has_fov=1, fov_deg=20.0 + 2.0 * k) for k in range(9)]
    outs = [np.zeros((360, 640, 4), np.uint8) for _ in settings]
    tickets = [two.RenderBegin(obj, s, o) for s, o in zip(settings, outs)]
    for t in tickets:
        two.RenderEnd(t)
    two.close()
    one = rt_mod.RayTracer(devices=[0])
    for s, o in zip(settings, outs):
        assert (one.RenderAsync(obj, s).pixels == o).all()
    one.close()


# ---- wavefront (one launch pair per depth) and fused tail (k_tail) are two schedules of the same arithmetic ---------------------
@pytest.mark.parametrize("tail_max", ["0", "1073741824", "20000"])
def test_wavefront_and_tail_schedules_agree_with_oracle(samples, monkeypatch, tail_max):
    monkeypatch.setenv("RTB_TAIL_MAX", tail_max)  # 0: pure wavefront; 2^30: every path in k_tail from depth 0; 20000: switch mid-frame
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        rt = rt_mod.RayTracer(bvh_mode=mode)
        for name, kw in (("test_scene_1", dict()), ("test_scene_2", dict(soft_shadows=1, light_size=3.0, glossy=1, roughness=0.05)), ("eval_scene", dict(aa_samples=2))):
            obj, osc, _ = samples[name]
            p = params(320, 240, 6, **kw)
            tex = rt.RenderAsync(obj, p)
            ref = osc.render(p)
            within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"{name} tail_max={tail_max} mode={mode}")
            s, c = rt.stats(), ref["counters"]
            if mode == abi.RTB_BVH_REFERENCE:
                assert same == 1.0
                assert (s.rays_primary, s.rays_continuation, s.rays_shadow, s.paths_hit_primary) == (c.rays_primary, c.rays_continuation, c.rays_shadow, c.primary_hits)
            assert s.reserved[0] == 0
        rt.close()


def test_shared_memory_staged_traversal_matches(samples, monkeypatch):
    """RTB_SMEM=1: k_traverse works out of a per-block shared-memory copy of nodes + triangles (small scenes only)."""
    monkeypatch.setenv("RTB_SMEM", "1")
    monkeypatch.setenv("RTB_TAIL_MAX", "0")
    for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
        rt = rt_mod.RayTracer(bvh_mode=mode)
        for name in synth.SAMPLE_SCENES:
            obj, osc, _ = samples[name]
            p = params(400, 300, 6)
            tex = rt.RenderAsync(obj, p)
            ref = osc.render(p)
            within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"smem {name} mode {mode}")
            if mode == abi.RTB_BVH_REFERENCE:
                assert same == 1.0
        rt.close()


# ---- analytic primitive mode: spheres / boxes with the semantics of the reference's HittableObjects.cs ----------------
@pytest.mark.parametrize("mode", [abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH])
@pytest.mark.parametrize("scene_name", ["test_scene_1", "eval_scene", "grid4", "grid_rot"])
def test_analytic_primitive_mode_matches_oracle(oracle, samples, mode, scene_name):
    if scene_name == "grid4":
        obj = synth.sphere_grid_scene(4)
    elif scene_name == "grid_rot":  # rotated, non-uniformly scaled spheres and boxes: exercises the inverse-transpose normals
        obj = synth.sphere_grid_scene(3)
        T = scene_mod.TransformElement
        for k, tr in enumerate(obj.Transformations[3:12]):
            tr.Elements += [T.RotationX(20.0 * k), T.RotationZ(35.0), T.Scale((1.0 + 0.2 * k, 0.7, 1.3))]
        obj.Transformations.append(scene_mod.CompositeTransformation([T.Translation((0.0, 0.0, 6.0)), T.RotationY(30.0), T.Scale((4.0, 2.0, 1.0))]))
        obj.Boxes.append(scene_mod.BoxDescription(len(obj.Transformations) - 1, 1))
    else:
        obj = samples[scene_name][0]
    osc, holder = oracle_scene(oracle, obj)
    osc.set_primitive_mode(abi.RTB_PRIM_ANALYTIC)
    rt = rt_mod.RayTracer(bvh_mode=mode, primitive_mode=abi.RTB_PRIM_ANALYTIC)
    for kw in (dict(), dict(is_orthographic=1), dict(debug_mode=2)):
        p = params(400, 300, 8, **kw)
        tex = rt.RenderAsync(obj, p)
        ref = osc.render(p, want_aux=True)
        within, same, worst = assert_rgb_parity(tex.pixels, ref["rgba8"], f"analytic {scene_name} mode {mode} {kw}")
        if not kw:
            s, c = rt.stats(), ref["counters"]
            assert s.n_triangles == osc.n_primitives
            assert abs((s.rays_primary + s.rays_continuation + s.rays_shadow) - c.rays) <= 0.001 * c.rays
            prim, t, mat = rt.primary_hits(obj, p)
            assert (t.view(np.uint32) == ref["t"].view(np.uint32)).mean() >= 0.9999
            assert (prim == ref["prim"]).mean() >= 0.999 and (mat == ref["mat"]).mean() >= 0.999
            assert s.reserved[0] == 0
    rt.close()


# ---- pipelined host API ------------------------------------------------------------------------------------------------------------
def test_render_begin_end_pipeline_matches_blocking(samples, monkeypatch):
    """Frames in flight (two lanes, two device frame buffers) must land in their own host buffers, bit-identical to rtb_render."""
    monkeypatch.setenv("RTB_CHUNK_SLOTS", "60000")  # several chunks per frame, alternating lanes
    obj = samples["test_scene_2"][0]
    rt = rt_mod.RayTracer()
    settings = [params(320, 200, 4, has_fov=1, fov_deg=20.0 + 3.0 * k) for k in range(12)]
    want = [rt.RenderAsync(obj, p).pixels for p in settings]
    outs = [np.zeros((200, 320, 4), np.uint8) for _ in settings]
    tickets = [rt.RenderBegin(obj, p, o) for p, o in zip(settings, outs)]  # 12 > 8: the ticket ring recycles
    for t in tickets:
        rt.RenderEnd(t)
    for k in range(len(settings)):
        assert (outs[k] == want[k]).all(), k
    assert len({w.tobytes() for w in want}) > 6  # the frames really differ
    with pytest.raises(rt_mod.RtbError):
        rt.RenderEnd(10 ** 6)
    rt.close()


def test_rotation_sweep_matches_frame_by_frame_renders(samples, oracle):
    """GifGenerator.GenerateRotationFrames (GifGenerator.cs:40-72): 36 frames, CameraRotationOverride.z = 0, 10, ..., 350."""
    gif_mod = importlib.import_module("cosig-raytracing_b200.gif_generator")
    obj, osc, _ = samples["test_scene_1"]
    rt = rt_mod.RayTracer()
    base = scene_mod.RenderSettings(ResolutionOverride=(160, 120), MaxDepth=3, CameraPositionOverride=(0.0, 0.0, 0.0), CameraRotationOverride=(-60.0, 0.0, 0.0))
    seen = []
    frames = gif_mod.GifGenerator(rt, obj).GenerateRotationFrames(base, progress=lambda v, msg: seen.append(msg))
    assert len(frames) == 36 and len(seen) == 36 and seen[-1].startswith("Rendering frame 36/36")
    for k in (0, 7, 35):
        st = scene_mod.RenderSettings(ResolutionOverride=(160, 120), MaxDepth=3, CameraPositionOverride=(0.0, 0.0, 0.0), CameraRotationOverride=(-60.0, 0.0, 10.0 * k))
        assert (frames[k].pixels == osc.render(st.to_params())["rgba8"]).all(), k
    assert len({f.pixels.tobytes() for f in frames}) >= 30
    rt.close()


def test_c5_shape_8k_16spp_many_chunks(tracers, samples):
    """BASELINE config C5's shape (7680x4320, 16 spp, depth 6: 531 M pixel-samples in 32 chunks over the lanes) on the sample
    scene; the oracle checks 16 rows spread over the frame, which cross chunk and lane boundaries."""
    obj, osc, _ = samples["test_scene_1"]
    p = params(7680, 4320, 6, 16)
    rt = tracers(abi.RTB_BVH_REFERENCE)
    tex = rt.RenderAsync(obj, p).pixels
    st = rt.stats()
    assert st.chunks >= 30 and st.rays_primary == 7680 * 4320 * 16 and st.reserved[0] == 0
    rows = np.arange(7, 4320, 270)
    ref = osc.render(p, rows=(7, -1, 270))
    assert (tex[rows] == ref["rgba8"][rows]).all()


def test_exported_frame_is_never_moved_under_its_importers(samples):
    """once rtb_frame_export has handed out an IPC handle, a render that would need a larger internal frame and
    rtb_clear_target must fail instead of freeing memory peers may still store into; re-uploads reuse pooled memory."""
    obj = samples["test_scene_1"][0]
    rt = rt_mod.RayTracer()
    ptr, handle = rt.frame_export(64 * 48 * 4)
    assert rt.RenderAsync(obj, params(64, 48, 2)) is not None          # fits the exported buffer
    with pytest.raises(rt_mod.RtbError):
        rt.RenderAsync(obj, params(128, 96, 2))                          # would have to reallocate it
    with pytest.raises(rt_mod.RtbError):
        rt.ClearRenderTarget()
    with pytest.raises(rt_mod.RtbError):
        rt.frame_export(128 * 96 * 4)
    ptr2, _ = rt.frame_export(64 * 48 * 4)
    assert ptr2 == ptr
    rt.close()
    rt = rt_mod.RayTracer()
    a = rt.RenderAsync(obj, params(160, 120, 3)).pixels
    for _ in range(3):                                                   # invalidate + upload cycles: pooled memory, same frame
        rt.InvalidateBVHCache()
        assert (rt.RenderAsync(obj, params(160, 120, 3)).pixels == a).all()
    rt.ReleaseBuffers()
    assert (rt.RenderAsync(obj, params(160, 120, 3)).pixels == a).all()
    rt.close()
