import os, sys
sys.path.insert(0, "/root/repo")
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, _lib
name = sys.argv[1]
desc = scenes.WORKLOADS[name]()
s = Scene(desc)
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
integ.SampleF32(2, flags=_lib.SAMPLE_COUNT_OWN_TREE)
st = integ.stats
print(name, "records/ray %.3f %.3f tris/ray %.3f %.3f" % (st["nodes"][0]/st["closest_rays"], st["nodes"][1]/st["shadow_rays"], st["tris"][0]/st["closest_rays"], st["tris"][1]/st["shadow_rays"]))
