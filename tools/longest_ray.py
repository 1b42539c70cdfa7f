"""Diagnostic (library built with -DRTB_RAY_STATS=1): node visits of the longest ray of a frame, per max depth, on C4 / C3."""
import importlib, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from util import abi, params, synth
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
for name, scene in (("c4 heightfield", synth.heightfield_scene(1000, 500)), ("c3 sphere grid", synth.sphere_grid_scene(16))):
    rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
    for depth, diffuse in ((1, 0), (1, 1), (2, 1), (6, 1)):
        p = params(3840, 2160, depth, enable_diffuse=diffuse)
        rt.RenderToTexture(scene, p)
        st = rt.stats()
        rays = st.rays_primary + st.rays_continuation + st.rays_shadow
        print(f"{name}: depth {depth} shadows {diffuse}: rays {rays}, mean node records / ray {st.reserved[1] / rays:.1f}, longest ray {st.reserved[3]} node records", flush=True)
    rt.close()
