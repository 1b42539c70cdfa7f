"""Small renders through every kernel variant.  Written for compute-sanitizer (memcheck; keep it tiny: the tool slows kernels ~50x);
where the sanitizer is not available it still is a functional sweep: every optional traversal kernel must reproduce the default
kernel's frame and primary hits, and pipelined frames must equal blocking ones."""
import importlib, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util import abi, params, scene_mod, synth, tiny_scene
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")

def run(tag, env, mode, prim, obj, p):
    for k in ("RTB_TAIL_MAX", "RTB_SMEM", "RTB_CHUNK_SLOTS", "RTB_LANES", "RTB_WIDE", "RTB_POOL", "RTB_PACKET_CLOSEST", "RTB_PACKET_SHADOW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    rt = rt_mod.RayTracer(bvh_mode=mode, primitive_mode=prim)
    a = rt.RenderAsync(obj, p).pixels
    out = [np.zeros_like(a) for _ in range(3)]
    ts = [rt.RenderBegin(obj, p, o) for o in out]
    for t in ts:
        rt.RenderEnd(t)
    assert all((o == a).all() for o in out), tag
    hits = rt.primary_hits(obj, p)
    st = rt.stats()
    assert st.reserved[0] == 0, (tag, "traversal stack overflow")
    rt.close()
    print(tag, "ok", st.rays_primary, st.rays_continuation, st.rays_shadow, flush=True)
    return a, hits

s1 = synth.sample_scene("test_scene_1")
hf = synth.heightfield_scene(40, 20)
for mode in (abi.RTB_BVH_REFERENCE, abi.RTB_BVH_LBVH):
    run(f"wavefront mode{mode}", {"RTB_TAIL_MAX": "0", "RTB_SMEM": "0"}, mode, 0, s1, params(96, 64, 4, 2))
    run(f"tail+smem mode{mode}", {"RTB_TAIL_MAX": "1000000", "RTB_SMEM": "1"}, mode, 0, s1, params(96, 64, 4))
    run(f"smem wavefront chunks mode{mode}", {"RTB_TAIL_MAX": "0", "RTB_SMEM": "1", "RTB_CHUNK_SLOTS": "2048"}, mode, 0, s1, params(96, 64, 3))
    run(f"analytic mode{mode}", {"RTB_TAIL_MAX": "2000"}, mode, 1, s1, params(96, 64, 4, soft_shadows=1, light_size=2.0, glossy=1, roughness=0.1))
    run(f"heightfield mode{mode}", {}, mode, 0, hf, params(128, 72, 6))
    run(f"tiny mode{mode}", {}, mode, 0, tiny_scene(1), params(33, 17, 2, debug_mode=2))
# round 2: the optional traversal kernels (packets on both flavours, 8-wide quantised records, the regrouping pool), each against
# the default kernel's frame, and the hand-written radix sort / scan / dense emit on a scene large enough for several tiles
ref_frame, ref_hits = run("default lbvh", {"RTB_SMEM": "0"}, abi.RTB_BVH_LBVH, 0, hf, params(128, 72, 6))
for tag, env, mode in (("packets lbvh", {"RTB_PACKET_CLOSEST": "16", "RTB_PACKET_SHADOW": "16", "RTB_SMEM": "0"}, abi.RTB_BVH_LBVH),
                       ("packets reference", {"RTB_PACKET_CLOSEST": "1", "RTB_PACKET_SHADOW": "0"}, abi.RTB_BVH_REFERENCE),
                       ("wide records", {"RTB_WIDE": "1", "RTB_SMEM": "0"}, abi.RTB_BVH_LBVH),
                       ("wide records smem", {"RTB_WIDE": "1", "RTB_SMEM": "1"}, abi.RTB_BVH_LBVH),
                       ("pool", {"RTB_POOL": "1", "RTB_SMEM": "0"}, abi.RTB_BVH_LBVH)):
    frame, hits = run(tag, env, mode, 0, hf, params(128, 72, 6))
    if mode == abi.RTB_BVH_LBVH:  # order-independent closest hit (closer_hit's tie rule): identical t bits and ids, identical frame
        assert (hits[1].view(np.uint32) == ref_hits[1].view(np.uint32)).all() and (hits[0] == ref_hits[0]).all(), tag
        assert (frame == ref_frame).mean() > 0.9999, tag
    run(tag + " analytic", env, mode, 1, s1, params(96, 64, 4))
for k in ("RTB_WIDE", "RTB_POOL", "RTB_PACKET_CLOSEST", "RTB_PACKET_SHADOW", "RTB_SMEM"):
    os.environ.pop(k, None)
big = synth.heightfield_scene(150, 60)   # 18 000 triangles: five sort tiles, two scan levels of the digit table
run("own sort, 5 tiles", {}, abi.RTB_BVH_LBVH, 0, big, params(96, 54, 3))
# group ring of one rank and the external-memory path are exercised by the GPU tests (they need a second process / cuda-python)
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
rt.group_create(0, 1, 96 * 64 * 4, 2)
outs = [np.zeros((64, 96, 4), np.uint8) for _ in range(4)]
for t in [rt.GroupRenderBegin(s1, params(96, 64, 3), o) for o in outs]:
    rt.GroupRenderEnd(t)
assert all((o == outs[0]).all() for o in outs)
rt.group_create_host(0, 1, 96 * 64 * 4, 2)   # the host ring of one rank: same frames, out of the shared mapping
for t in [rt.GroupRenderBegin(s1, params(96, 64, 3)) for _ in range(2)]:
    rt.GroupRenderEnd(t)
    assert (rt.group_frame(t, 64, 96) == outs[0]).all()
rt.close()
print("group of one ok", flush=True)
# GIF sweep: palette kernel (vector and scalar paths), indexed pipelined readback, fused rotation call
gif = importlib.import_module("cosig-raytracing_b200.gif_generator")
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
for (h, w) in ((24, 32), (17, 19)):
    frame = np.random.RandomState(h).randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    out = np.zeros((h, w), np.uint8)
    rt._check(abi.load().rtb_gif_index_frame(rt._ctx, frame.ctypes.data, w, h, out.ctypes.data))
import tempfile
gif.GifGenerator(rt, s1).RenderRotationGif(scene_mod.RenderSettings(ResolutionOverride=(48, 32), MaxDepth=2), os.path.join(tempfile.mkdtemp(), "s.gif"), totalFrames=6, stepDeg=60.0)
rt.close()
print("gif ok", flush=True)
empty = scene_mod.ObjectData(); synth._sample_camera_and_light(empty)
run("empty", {}, abi.RTB_BVH_LBVH, 0, empty, params(40, 24, 3))
print("sanitize run complete")
