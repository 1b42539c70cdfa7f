import sys, importlib
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np
from util import *
from oracle import oracle_py as O
rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
obj = synth.sample_scene("test_scene_1")
osc, holder = oracle_scene(O, obj)
rt = rt_mod.RayTracer()
rt.RenderToTexture(obj, params(16,16,1))
vn, mat = rt.triangles()
ovn, omat, _ = osc.triangles()
bad = np.argwhere(vn.view(np.uint32) != ovn.view(np.uint32))
print(len(bad), bad[:20])
for i,j in bad[:10]:
    print(i, j, vn[i,j].hex(), ovn[i,j].hex(), vn[i], ovn[i])
