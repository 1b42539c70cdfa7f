"""Diagnostic (library built with -DRTB_RAY_STATS=1, RTB_TIMELINE=1): per k_traverse launch, when the queue ran dry and when the
last warp finished — full C4 frame and one rank's share of an 8-way split."""
import importlib, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ["RTB_TIMELINE"] = "1"; os.environ["RTB_LANES"] = "1"
from util import abi, params, synth
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
scene = synth.heightfield_scene(1000, 500)
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
for world in (1, 8):
    p = params(3840, 2160, 6)
    if world > 1:
        p.band_rank, p.band_world, p.band_rows = 0, world, 8
    for _ in range(2):
        rt.RenderToTexture(scene, p)
    print(f"--- band_world {world}", file=sys.stderr, flush=True)
    st = rt.stats()
    print(f"band_world {world}: rays {st.rays_primary + st.rays_continuation + st.rays_shadow}, device ms {st.ms_render_device:.3f}, longest ray {st.reserved[3]}", file=sys.stderr, flush=True)
rt.close()
