import sys, importlib, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from util import *
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
os.environ["RTB_TAIL_MAX"] = "0"
s = synth.heightfield_scene()
rt = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_LBVH)
for name, p in (("primary only", params(3840,2160,1,enable_diffuse=0)), ("depth1+shadow", params(3840,2160,1)), ("depth6", params(3840,2160,6))):
    rt.RenderToTexture(s, p); rt.RenderToTexture(s, p)
    st = rt.stats()
    steps = st.reserved[1] + st.reserved[2]
    print(f"{name}: lane-steps used {steps/1e6:.1f}M, occupied by batches {st.reserved[3]/1e6:.1f}M -> batch fill {steps/max(1,st.reserved[3]):.3f}  ms {st.ms_render_device:.3f}")
