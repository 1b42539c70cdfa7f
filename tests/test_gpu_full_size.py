"""GPU parity at BASELINE.json's FULL sizes (C3, C4, C5) and for the process-per-GPU gather, against the CPU oracle.

The oracle reaches these sizes through its checker-only leaf accelerator (oracle.cpp: LeafAccel — same results as the linear leaf
scan, tests/test_host_cpu.py proves it bit for bit), so C3 is checked on every 16th row of the 4K frame instead of one row.
Bars as everywhere: t bits / ids exact where the closest hit is unique, RGB within 1/255 per channel on >= 99.9 % of pixels.
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from util import abi, assert_rgb_parity, oracle_scene, params, rgb_agreement, scene_mod, synth

pytestmark = pytest.mark.gpu

rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def c4(oracle):
    obj = synth.heightfield_scene()
    osc, holder = oracle_scene(oracle, obj)
    return obj, osc, holder


# ---- C3: glass / mirror sphere grid, 3840x2160, depth 16 ---------------------------------------------------------------------------
def test_c3_full_size_lbvh_against_oracle(oracle):
    """BASELINE config C3 at full size through the GPU-built LBVH; the oracle renders every 16th row (135 rows, ~1.5 M rays).  The
    reference-shape BVH degenerates on this scene (two leaves of 98 310 triangles), which is exactly what the leaf accelerator
    is for."""
    obj = synth.sphere_grid_scene(16)
    osc, holder = oracle_scene(oracle, obj)
    assert osc.max_leaf > 90000 and osc.n_accelerated_leaves >= 2
    p = params(3840, 2160, 16)
    rows = np.arange(0, 2160, 16)
    ref = osc.render(p, rows=(0, -1, 16), want_aux=True)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        tex = rt.RenderAsync(obj, p).pixels
        st = rt.stats()
        assert st.n_triangles == 196620 and st.reserved[0] == 0
        assert_rgb_parity(tex[rows], ref["rgba8"][rows], "C3 LBVH vs oracle rows")
        prim, t, mat = rt.primary_hits(obj, p)
        same_t = t[rows].view(np.uint32) == ref["t"][rows].view(np.uint32)
        assert same_t.mean() >= 0.9999
        assert (mat[rows] == ref["mat"][rows]).mean() >= 0.9999
        # ray counts: every row the oracle rendered contributes its exact share; the whole frame is bounded by the per-row maximum
        assert st.rays_primary == 3840 * 2160


def test_c3_full_size_analytic_against_oracle(oracle):
    """C3 with analytic spheres / box: 257 primitives, every 16th row against the oracle's analytic mode."""
    obj = synth.sphere_grid_scene(16)
    osc, holder = oracle_scene(oracle, obj)
    osc.set_primitive_mode(abi.RTB_PRIM_ANALYTIC)
    p = params(3840, 2160, 16)
    rows = np.arange(0, 2160, 16)
    ref = osc.render(p, rows=(0, -1, 16))
    for mode in (abi.RTB_BVH_LBVH, abi.RTB_BVH_REFERENCE):
        with rt_mod.RayTracer(bvh_mode=mode, primitive_mode=abi.RTB_PRIM_ANALYTIC) as rt:
            tex = rt.RenderAsync(obj, p).pixels
            assert rt.stats().n_triangles == 257 and rt.stats().reserved[0] == 0
            assert_rgb_parity(tex[rows], ref["rgba8"][rows], f"C3 analytic mode {mode} vs oracle rows")


# ---- C4: 1 M triangles, the tie rule of the LBVH flavour --------------------------------------------------------------------------
def test_c4_lbvh_prim_ids_differ_only_at_exact_ties(c4):
    """'ids equal except exact ties' for C4 / LBVH, at 1 M triangles and 4K.  Every 16th row: t bits and material
    equal wherever the oracle has a unique closest hit; where the prim id differs, brute force over all 1 M triangles must find
    several triangles at exactly the closest t, the GPU's id among them (the LBVH flavour then keeps the smallest leaf index)."""
    obj, osc, _ = c4
    p = params(3840, 2160, 6)
    rows = np.arange(8, 2160, 16)
    ref = osc.render(p, rows=(8, -1, 16), want_aux=True)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        prim, t, mat = rt.primary_hits(obj, p)
        assert rt.stats().reserved[0] == 0
    same_t = t[rows].view(np.uint32) == ref["t"][rows].view(np.uint32)
    assert same_t.mean() >= 0.9999, f"t mismatch on {(~same_t).sum()} pixels"
    differ = np.argwhere((prim[rows] != ref["prim"][rows]) & same_t)
    assert len(differ) <= 0.002 * same_t.size
    for yi, x in differ[:60]:
        y = int(rows[yi])
        o, d = osc.primary_ray(p, int(x), y)
        tb, ids, n = osc.brute_closest(o, d, cap=64)
        assert n >= 2 and prim[y, x] in ids, f"pixel ({x},{y}): id {prim[y, x]} vs {ref['prim'][y, x]} without a tie"
    assert (mat[rows] == ref["mat"][rows])[same_t].mean() >= 0.999


def test_c4_binary_and_wide_records_agree(c4, monkeypatch):
    """The two node formats of the LBVH flavour (8-wide quantised, binary two-box) must give the same primary hits and frames: the
    closest hit does not depend on which boxes were opened, nor on the order (closer_hit's tie rule)."""
    obj, _, _ = c4
    p = params(1920, 1080, 6)
    out = {}
    for wide in ("1", "0"):
        monkeypatch.setenv("RTB_WIDE", wide)
        with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
            tex = rt.RenderAsync(obj, p).pixels
            out[wide] = (tex,) + rt.primary_hits(obj, p)
            assert rt.stats().reserved[0] == 0
    a, b = out["1"], out["0"]
    assert (a[2].view(np.uint32) == b[2].view(np.uint32)).all(), "primary t bits differ between node formats"
    assert (a[1] == b[1]).all(), "primary prim ids differ between node formats"
    assert (a[0] == b[0]).mean() >= 0.99999


# ---- C5: the C4 scene at 7680x4320, 16 spp ----------------------------------------------------------------------------------------
def test_c5_on_the_c4_scene_sampled_rows(c4):
    """BASELINE config C5 proper (1 M triangles, 8K, 16 spp, depth 6: 531 M pixel-samples in 32 chunks): the oracle renders 9 rows
    spread over the frame (1.1 M pixel-samples, ~2 M rays)."""
    obj, osc, _ = c4
    p = params(7680, 4320, 6, 16)
    rows = np.arange(150, 4320, 500)
    ref = osc.render(p, rows=(150, -1, 500))
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        tex = rt.RenderAsync(obj, p).pixels
        st = rt.stats()
        assert st.chunks >= 30 and st.rays_primary == 7680 * 4320 * 16 and st.reserved[0] == 0
    assert_rgb_parity(tex[rows], ref["rgba8"][rows], "C5 on the C4 scene, LBVH vs oracle rows")


# ---- camera far from the scene (the LBVH box tests must stay conservative whatever the distance) -----------------------------------
@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("distance", [74.0, 3000.0, 60000.0])
def test_far_camera_lbvh_matches_reference_mode(distance, wide, monkeypatch):
    """the LBVH box tests are evaluated as fma(bound, 1/d, -origin/d), which rounds by ~1e-7 |origin / d|: build-time
    padding alone only covered cameras within a few scene radii.  Both node formats now widen per ray (the binary records by
    2^-21 |origin / d| per axis, the wide records by 2^-20 of their plane terms), so a camera 1000 scene radii away must still see
    every hit the reference-shape tree sees."""
    monkeypatch.setenv("RTB_WIDE", wide)
    obj = synth.heightfield_scene(100, 50)
    T = scene_mod.TransformElement
    obj.Transformations[1] = scene_mod.CompositeTransformation([T.Translation((0, 0, -distance)), T.RotationX(-60.0), T.RotationZ(45.0)])
    fov = float(np.degrees(2.0 * np.arctan(np.tan(np.radians(15.0)) * 74.0 / distance)))
    p = params(640, 360, 4, has_fov=1, fov_deg=fov)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_REFERENCE) as rt:
        ref_prim, ref_t, _ = rt.primary_hits(obj, p)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        prim, t, _ = rt.primary_hits(obj, p)
    hit_ref, hit = ref_prim >= 0, prim >= 0
    assert hit_ref.mean() > 0.2
    assert (hit_ref & ~hit).sum() == 0, f"{(hit_ref & ~hit).sum()} hits of the reference-shape tree are lost by the LBVH at distance {distance}"
    assert (t.view(np.uint32) == ref_t.view(np.uint32)).mean() >= (0.9999 if distance < 100.0 else 0.999)


# ---- process-per-GPU gather on ONE GPU: two processes, CUDA IPC, peer stores from k_resolve ------------------------------------------
_CHILD = r"""
import importlib, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from util import abi, params, synth
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
handle = bytes.fromhex(sys.stdin.readline().strip())
obj = synth.sample_scene("test_scene_2")
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
dst = rt.frame_import(handle)
p = params({w}, {h}, 5, 2, band_rank=1, band_world=2, band_rows=8)
rt.RenderToTexture(obj, p, dst, {w} * {h} * 4, sync=True)
print("stored", flush=True)
sys.stdin.readline()   # keep the mapping alive until the parent has read the frame
rt.close()
"""


def test_two_processes_gather_into_one_frame_over_ipc(samples_scene2):
    """The N > 1 form of bench.py on a single GPU: rank 0 (this process) exports its frame buffer, rank 1 (a second process on the
    same device) maps it through CUDA IPC and its k_resolve stores its bands straight into it; rank 0 renders its own bands, waits
    for rank 1 and reads the frame back.  Must equal the one-context frame bit for bit."""
    obj = samples_scene2
    w, h = 640, 360
    p_full = params(w, h, 5, 2)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as one:
        want = one.RenderAsync(obj, p_full).pixels
    rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
    try:
        ptr, handle = rt.frame_export(w * h * 4)
        child = subprocess.Popen([sys.executable, "-c", _CHILD.format(root=ROOT, tests=os.path.join(ROOT, "tests"), w=w, h=h)],
                                 stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
        child.stdin.write(handle.hex() + "\n"); child.stdin.flush()
        rt.RenderToTexture(obj, params(w, h, 5, 2, band_rank=0, band_world=2, band_rows=8), ptr, w * h * 4, sync=True)
        line = ""
        for _ in range(50):  # skip anything a library prints while loading
            line = child.stdout.readline()
            if not line or line.strip() == "stored":
                break
        assert line.strip() == "stored", f"child ended with {line!r} (exit code {child.poll()})"
        got = np.zeros((h, w, 4), np.uint8)
        rt.frame_read(got)
        child.stdin.write("done\n"); child.stdin.flush()
        assert child.wait(timeout=120) == 0
    finally:
        rt.close()
    assert (got == want).all(), f"{(got != want).any(axis=-1).sum()} pixels differ"


@pytest.fixture(scope="module")
def samples_scene2():
    return synth.sample_scene("test_scene_2")


# ---- zero-copy display path: render into memory owned by another API ---------------------------------------------------
def test_render_into_imported_external_memory(samples_scene2):
    """RenderToTexture without readback (RayTracer.cs:82-202): the frame lands in an allocation the library did not make.  A CUDA
    virtual-memory allocation exported as a POSIX file descriptor stands in for the Vulkan / D3D12 buffer Unity would export
    (vkGetMemoryFdKHR): rtb_external_import maps it (cudaImportExternalMemory), rtb_render_device's resolve kernel stores into it,
    and the 'graphics side' reads the pixels through its own mapping of the same memory."""
    try:
        from cuda.bindings import driver as cu
    except ImportError:
        from cuda import cuda as cu
    obj = samples_scene2
    w, h = 512, 288
    p = params(w, h, 4)

    def ok(res):
        err, rest = res[0], res[1:]
        assert int(err) == 0, f"driver API error {err}"
        return rest[0] if len(rest) == 1 else rest

    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        want = rt.RenderAsync(obj, p).pixels
        ok(cu.cuInit(0))
        dev = ok(cu.cuDeviceGet(0))
        ctx = ok(cu.cuDevicePrimaryCtxRetain(dev))
        ok(cu.cuCtxSetCurrent(ctx))
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = 0
        prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        gran = ok(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
        size = (w * h * 4 + gran - 1) // gran * gran
        handle = ok(cu.cuMemCreate(size, prop, 0))
        fd = ok(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
        va = ok(cu.cuMemAddressReserve(size, 0, 0, 0))
        ok(cu.cuMemMap(va, size, 0, handle, 0))
        acc = cu.CUmemAccessDesc()
        acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = 0
        acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        ok(cu.cuMemSetAccess(va, size, [acc], 1))
        ok(cu.cuMemsetD8(va, 0, size))
        ptr = rt.external_import(abi.RTB_EXT_OPAQUE_FD, int(fd), size)
        assert ptr
        rt.RenderToTexture(obj, p, ptr, w * h * 4, sync=True)
        got = np.zeros((h, w, 4), np.uint8)
        ok(cu.cuMemcpyDtoH(got.ctypes.data, va, got.nbytes))
        assert (got == want).all(), f"{(got != want).any(axis=-1).sum()} pixels differ"
        # a second frame with other settings into the same mapping (the realtime loop reuses its target, RayTracer.cs:126-132)
        p2 = params(w, h, 2, has_fov=1, fov_deg=20.0)
        want2 = rt.RenderAsync(obj, p2).pixels
        rt.RenderToTexture(obj, p2, ptr, w * h * 4, sync=True)
        ok(cu.cuMemcpyDtoH(got.ctypes.data, va, got.nbytes))
        assert (got == want2).all()
        rt.external_release(ptr)
        with pytest.raises(rt_mod.RtbError):
            rt.external_release(ptr)
        ok(cu.cuMemUnmap(va, size))
        ok(cu.cuMemAddressFree(va, size))
        ok(cu.cuMemRelease(handle))


# ---- the process-per-GPU frame ring inside the library (rtb_group_*) -------------------------------------------------------------------
_GROUP_CHILD = r"""
import importlib, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from util import abi, params, synth
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
handle = bytes.fromhex(sys.stdin.readline().strip())
obj = synth.sample_scene("test_scene_2")
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
rt.group_create(1, 2, {w} * {h} * 4, {nbuf}, handle)
print("joined", flush=True)
tickets = [rt.GroupRenderBegin(obj, params({w}, {h}, 4, 1, has_fov=1, fov_deg=20.0 + 2.0 * k)) for k in range({frames})]
for t in tickets:
    rt.GroupRenderEnd(t)
print("stored", flush=True)
sys.stdin.readline()
rt.close()
"""


def test_group_frame_ring_two_processes_one_gpu(samples_scene2):
    """rtb_group_*: two ranks (two processes; here on the same GPU) render the bands of a stream of frames into rank 0's ring of
    frame buffers and rank 0 reads them back, with more frames in flight than buffers.  Every frame must equal the one-context
    render bit for bit, in order."""
    obj = samples_scene2
    w, h, frames, nbuf = 400, 240, 7, 2
    settings = [params(w, h, 4, 1, has_fov=1, fov_deg=20.0 + 2.0 * k) for k in range(frames)]
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as one:
        want = [one.RenderAsync(obj, p).pixels for p in settings]
    rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
    child = None
    try:
        handle = rt.group_create(0, 2, w * h * 4, nbuf)
        child = subprocess.Popen([sys.executable, "-c", _GROUP_CHILD.format(root=ROOT, tests=os.path.join(ROOT, "tests"), w=w, h=h, nbuf=nbuf, frames=frames)],
                                 stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
        child.stdin.write(handle.hex() + "\n"); child.stdin.flush()

        def expect(word):
            line = ""
            for _ in range(50):
                line = child.stdout.readline()
                if not line or line.strip() == word:
                    break
            assert line.strip() == word, f"child ended with {line!r} instead of {word!r} (exit code {child.poll()})"

        expect("joined")
        outs = [np.zeros((h, w, 4), np.uint8) for _ in settings]
        tickets = [rt.GroupRenderBegin(obj, p, o) for p, o in zip(settings, outs)]
        for t in tickets:
            rt.GroupRenderEnd(t)
        expect("stored")
        for k, (o, wnt) in enumerate(zip(outs, want)):
            assert (o == wnt).all(), f"frame {k}: {(o != wnt).any(axis=-1).sum()} pixels differ"
        child.stdin.write("done\n"); child.stdin.flush()
        assert child.wait(timeout=120) == 0
    finally:
        if child is not None and child.poll() is None:
            child.kill()
        rt.close()


def test_group_of_one_equals_render_begin(samples_scene2):
    """world = 1: the ring degenerates to rtb_render_begin / rtb_render_end (and a missing peer is an error, not a hang)."""
    obj = samples_scene2
    p = params(320, 200, 3)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        want = rt.RenderAsync(obj, p).pixels
        rt.group_create(0, 1, 320 * 200 * 4, 2)
        outs = [np.zeros((200, 320, 4), np.uint8) for _ in range(5)]
        for t in [rt.GroupRenderBegin(obj, p, o) for o in outs]:
            rt.GroupRenderEnd(t)
        assert all((o == want).all() for o in outs)
        rt.group_destroy()
    import os as _os
    _os.environ["RTB_GROUP_TIMEOUT_MS"] = "300"
    try:
        with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
            rt.group_create(0, 2, 320 * 200 * 4, 2)  # rank 1 never joins
            out = np.zeros((200, 320, 4), np.uint8)
            t = rt.GroupRenderBegin(obj, p, out)
            with pytest.raises(rt_mod.RtbError):
                rt.GroupRenderEnd(t)
    finally:
        del _os.environ["RTB_GROUP_TIMEOUT_MS"]


# ---- the host ring: every rank copies its own bands into shared page-locked memory (rtb_group_create_host) ------------------------------
_HOST_RING_CHILD = r"""
import importlib, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from util import abi, params, synth
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
name = sys.stdin.readline().strip()
obj = synth.sample_scene("test_scene_2")
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
rt.group_create_host({rank}, {world}, {w} * {h} * 4, {nbuf}, name)
print("joined", flush=True)
tickets = []
for k in range({frames}):
    if k >= {nbuf}:
        rt.GroupRenderEnd(tickets[k - {nbuf}])
    tickets.append(rt.GroupRenderBegin(obj, params({w}, {h}, 4, 1, has_fov=1, fov_deg=20.0 + 2.0 * k, band_rows={band_rows})))
for t in tickets[-{nbuf}:]:
    rt.GroupRenderEnd(t)
print("stored", flush=True)
sys.stdin.readline()
rt.close()
"""


@pytest.mark.parametrize("world,h,band_rows", [(2, 240, 8), (3, 236, 16)])
def test_host_ring_processes_on_one_gpu(samples_scene2, world, h, band_rows):
    """rtb_group_create_host: `world` ranks (processes; here all on one GPU) render the bands of a stream of frames and each copies ITS
    rows into the frame's slot of a ring in shared page-locked host memory; rank 0 reads whole frames at rtb_group_frame.  More frames
    than buffers, so slots are reused under the begun-by-rank-0 rule.  236 rows in bands of 16 over 3 ranks: the last band is short (12
    rows) and the ranks own 5 / 5 / 5 bands.  Every frame must equal the one-context render bit for bit, in order."""
    obj = samples_scene2
    w, frames, nbuf = 400, 7, 2
    settings = [params(w, h, 4, 1, has_fov=1, fov_deg=20.0 + 2.0 * k, band_rows=band_rows) for k in range(frames)]
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as one:
        want = [one.RenderAsync(obj, params(w, h, 4, 1, has_fov=1, fov_deg=20.0 + 2.0 * k)).pixels for k in range(frames)]
    rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
    children = []
    try:
        name = rt.group_create_host(0, world, w * h * 4, nbuf)
        for r in range(1, world):
            c = subprocess.Popen([sys.executable, "-c", _HOST_RING_CHILD.format(root=ROOT, tests=os.path.join(ROOT, "tests"), rank=r, world=world, w=w, h=h,
                                                                                nbuf=nbuf, frames=frames, band_rows=band_rows)],
                                 stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
            c.stdin.write(name + "\n"); c.stdin.flush()
            children.append(c)

        def expect(c, word):
            line = ""
            for _ in range(50):
                line = c.stdout.readline()
                if not line or line.strip() == word:
                    break
            assert line.strip() == word, f"child ended with {line!r} instead of {word!r} (exit code {c.poll()})"

        for c in children:
            expect(c, "joined")
        got, tickets = [], []
        for k, p in enumerate(settings):
            if k >= nbuf:  # the frame leaves the ring when frame k is begun: take it out first
                rt.GroupRenderEnd(tickets[k - nbuf])
                got.append(rt.group_frame(tickets[k - nbuf], h, w).copy())
            tickets.append(rt.GroupRenderBegin(obj, p))
        for t in tickets[-nbuf:]:
            rt.GroupRenderEnd(t)
            got.append(rt.group_frame(t, h, w).copy())
        with pytest.raises(rt_mod.RtbError):
            rt.group_frame(tickets[0], h, w)  # long gone
        for c in children:
            expect(c, "stored")
        for k, (o, wnt) in enumerate(zip(got, want)):
            assert (o == wnt).all(), f"frame {k}: {(o != wnt).any(axis=-1).sum()} pixels differ"
        for c in children:
            c.stdin.write("done\n"); c.stdin.flush()
        for c in children:
            assert c.wait(timeout=120) == 0
    finally:
        for c in children:
            if c.poll() is None:
                c.kill()
        rt.close()


def test_host_ring_of_one_and_its_errors(samples_scene2):
    """world = 1: the host ring is rtb_render_begin / _end into library-owned memory.  A caller buffer is refused (the frame lives in the
    ring), rtb_render_begin is refused while the group owns the pipelined buffers, a missing peer is an error after the timeout."""
    obj = samples_scene2
    p = params(320, 200, 3)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        want = rt.RenderAsync(obj, p).pixels
        name = rt.group_create_host(0, 1, 320 * 200 * 4, 3)
        assert os.path.exists("/dev/shm" + name)
        for rounds in range(3):
            tickets = [rt.GroupRenderBegin(obj, p) for _ in range(3)]
            for t in tickets:
                rt.GroupRenderEnd(t)
                assert (rt.group_frame(t, 200, 320) == want).all()
        tickets = [rt.GroupRenderBegin(obj, p) for _ in range(3)]
        with pytest.raises(rt_mod.RtbError):   # a fourth frame would overwrite the first, which nobody has read yet
            rt.GroupRenderBegin(obj, p)
        rt.GroupRenderEnd(tickets[0])
        tickets.append(rt.GroupRenderBegin(obj, p))
        for t in tickets[1:]:
            rt.GroupRenderEnd(t)
            assert (rt.group_frame(t, 200, 320) == want).all()
        out = np.zeros((200, 320, 4), np.uint8)
        with pytest.raises(rt_mod.RtbError):
            rt.GroupRenderBegin(obj, p, out)
        with pytest.raises(rt_mod.RtbError):
            rt.RenderBegin(obj, p, out)
        with pytest.raises(rt_mod.RtbError):
            rt.group_create_host(0, 1, 320 * 200 * 4, 3, "no-leading-slash")
        rt.group_destroy()
        assert not os.path.exists("/dev/shm" + name)
        assert (rt.RenderAsync(obj, p).pixels == want).all()
        with pytest.raises(rt_mod.RtbError):
            rt.group_create_host(1, 2, 320 * 200 * 4, 3, "/rtb200-test-nobody-made-this")
    os.environ["RTB_GROUP_TIMEOUT_MS"] = "300"
    try:
        with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
            rt.group_create_host(0, 2, 320 * 200 * 4, 2)  # rank 1 never joins
            t = rt.GroupRenderBegin(obj, p)
            with pytest.raises(rt_mod.RtbError):
                rt.GroupRenderEnd(t)
    finally:
        del os.environ["RTB_GROUP_TIMEOUT_MS"]


# ---- every optional traversal kernel reproduces the default one ---------------------------------------------------------------------------
def test_optional_kernels_reproduce_the_default_kernel():
    """tools/sanitize_run.py: small renders through every kernel variant — wavefront / tail / shared-memory schedules in both BVH modes, the
    packet kernels on both flavours, the 8-wide quantised records (global and shared memory), the regrouping pool, analytic primitives
    under each, the hand-written sort on several tiles, a group of one, the GIF path — asserting identical primary t bits / ids and
    frames against the default LBVH kernel, pipelined frames equal to blocking ones, and no traversal-stack overflow."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_run.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sanitize run complete" in r.stdout
    for tag in ("packets lbvh ok", "packets reference ok", "wide records ok", "wide records smem ok", "pool ok", "own sort, 5 tiles ok", "group of one ok"):
        assert tag in r.stdout, tag


# ---- randomised parity sweep ----------------------------------------------------------------------------------------------------------------
def test_randomised_parity_sweep():
    """tools/parity_fuzz.py on 160 seeds (profiles/r2_parity_fuzz.log holds a 3000-seed run): random meshes / spheres / boxes under random
    composite transformations, random materials incl. out-of-range indices, random cameras and every render setting.  Reference-shape
    flavour: frame, primary ids / t bits / materials and ray counters identical to the oracle.  LBVH flavour: every primary hit that differs
    from the oracle's traversal is the brute-force closest hit, frames equal the oracle's exact-closest mode; 8-wide records equal the binary
    ones; analytic frames identical up to equal-t ties between coincident primitives."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_fuzz.py"), "--seeds", "160"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "160 seeds" in r.stdout and " 0 failures" in r.stdout


# ---- structure of the GPU-built hierarchies ---------------------------------------------------------------------------------------------
def _leaf_range(ref):
    code = ~int(ref)
    return code >> 3, (code & 7) + 1


def _check_tree(nodes_u32, nodes_f32, perm, tri_min, tri_max, wide):
    """Walks the records from the root: every triangle is referenced by exactly one leaf, every child box contains the triangles
    (leaf) or the child boxes (inner node) below it, every record is reached exactly once.  Returns (records reached, max depth)."""
    n_tris = len(perm)
    covered = np.zeros(n_tris, np.int32)
    reached = np.zeros(len(nodes_u32), np.int32)

    def children(i):
        """[(ref, box_min[3], box_max[3])] of record i."""
        if not wide:
            f, u = nodes_f32[i], nodes_u32[i]
            return [(np.int32(u[3]), f[0:3], f[4:7]), (np.int32(u[7]), f[8:11], f[12:15])]
        f, u = nodes_f32[i], nodes_u32[i]
        p = f[0:3].astype(np.float64)
        hdr = int(u[3])
        cell = np.array([2.0 ** (((hdr >> (8 * a)) & 255) - 127 - 15) for a in range(3)])
        valid = hdr >> 24
        refs = np.concatenate([u[4:8], u[20:24]]).astype(np.uint32).view(np.int32)
        q = u[8:20].astype(np.uint32).view(np.uint8).reshape(6, 8)  # lo.x lo.y lo.z hi.x hi.y hi.z, one byte per slot
        out = []
        for s in range(8):
            if not (valid >> s) & 1:
                assert refs[s] == np.int32(-2 ** 31), "an empty slot must carry the DONE reference"
                continue
            lo = p + q[0:3, s] * cell
            hi = p + q[3:6, s] * cell
            out.append((refs[s], lo, hi))
        return out

    def grid_cell(i):
        hdr = int(nodes_u32[i][3])
        return max(2.0 ** (((hdr >> (8 * a)) & 255) - 127 - 15) for a in range(3))

    stack = [(0, None, None, 1)]
    max_depth = 0
    while stack:
        i, bmin, bmax, depth = stack.pop()
        max_depth = max(max_depth, depth)
        reached[i] += 1
        for ref, lo, hi in children(i):
            lo64, hi64 = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
            assert (lo64 <= hi64).all()
            if bmin is not None:  # nested in the box its parent holds for this record; a wide record rounds its children outward on ITS grid, so they may stick out by one of its cells
                slack = 1e-4 * (1.0 + np.abs(bmax - bmin).max()) + (grid_cell(i) if wide else 0.0)
                assert (lo64 >= bmin - slack).all() and (hi64 <= bmax + slack).all(), f"record {i}: child box leaves its parent's box"
            if ref < 0:
                first, count = _leaf_range(ref)
                assert 1 <= count <= 4 and first + count <= n_tris
                ids = perm[first:first + count]
                covered[first:first + count] += 1
                tol = 1e-5 * (1.0 + np.abs(hi64).max())  # float32 rounding of (x - p) / cell at build time; the traversal widens by more
                assert (tri_min[ids] >= lo64 - tol).all() and (tri_max[ids] <= hi64 + tol).all(), f"record {i}: a leaf's triangle sticks out of its box"
            else:
                stack.append((int(ref), lo64, hi64, depth + 1))
    assert (covered == 1).all(), f"{(covered != 1).sum()} triangles are not referenced exactly once"
    assert (reached == 1).all(), f"{(reached != 1).sum()} records are not reached exactly once (the array must be dense)"
    return int(reached.sum()), max_depth


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("scene_name", ["heightfield", "test_scene_2"])
def test_gpu_built_hierarchy_is_a_valid_bvh(scene_name, wide, monkeypatch):
    """K2: the tree the GPU builds — Morton keys, the hand-written radix sort, Karras hierarchy, refit, and either the dense binary
    two-box records or the 8-wide quantised records — bounds every triangle, references each exactly once, and has no dead records."""
    monkeypatch.setenv("RTB_WIDE", wide)
    obj = synth.heightfield_scene(200, 100) if scene_name == "heightfield" else synth.sample_scene(scene_name)
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH) as rt:
        rt.RenderAsync(obj, params(64, 48, 2))
        nodes, perm = rt.bvh()
        vn, _ = rt.triangles()
    assert nodes.shape[1] == (24 if wide == "1" else 16)
    assert sorted(perm.tolist()) == list(range(len(perm))), "perm is not a permutation of the triangles"
    v = vn[:, :9].reshape(-1, 3, 3).astype(np.float64)
    reached, depth = _check_tree(nodes.view(np.uint32), nodes, perm, v.min(axis=1), v.max(axis=1), wide == "1")
    assert reached == len(nodes)
    # the wide tree is shallower and smaller than the binary one
    if wide == "1":
        assert len(nodes) < len(perm) / 8 and depth <= 16
