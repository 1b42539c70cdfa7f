import sys, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from util import *
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
s = synth.heightfield_scene()
for mode in (abi.RTB_BVH_LBVH, abi.RTB_BVH_REFERENCE):
    rt = rt_mod.RayTracer(bvh_mode=mode)
    for name, p in (("primary only", params(3840,2160,1,enable_diffuse=0)), ("depth1+shadow", params(3840,2160,1)), ("depth6", params(3840,2160,6))):
        rt.RenderToTexture(s, p)
        rt.RenderToTexture(s, p)
        st = rt.stats()
        rays = st.rays_primary+st.rays_continuation+st.rays_shadow
        print(f"mode {mode} {name}: rays {rays} (p {st.rays_primary} c {st.rays_continuation} s {st.rays_shadow}) hits {st.paths_hit_primary} nodes/ray {st.reserved[1]/rays:.2f} tris/ray {st.reserved[2]/rays:.2f}  ms {st.ms_render_device:.3f}  Mrays/s {rays/st.ms_render_device/1e3:.1f}  nodes/hitray {st.reserved[1]/max(1,st.paths_hit_primary):.1f}")
    rt.close()
