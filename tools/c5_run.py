import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32, _lib
desc = scenes.c5_soup(); bvh = Bvh.Build(desc.prims)
s = Scene(desc, bvh=bvh)
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
integ.SampleF32(1)
st = integ.stats
rays = st["closest_rays"] + st["shadow_rays"]
print(f"C5 1spp: {rays/st['ms_total']/1e3:.1f} Mrays/s extend_ms={st['ms_extend']:.2f} shadow_ms={st['ms_shadow']:.2f} launches={st['launches_extend']}")
if len(sys.argv) > 1:
    integ.SampleF32(1, flags=_lib.SAMPLE_COUNT_TRAVERSAL)
    c = integ.stats
    for k, r in ((0, c["closest_rays"]), (1, c["shadow_rays"])):
        print("class", k, "nodes/ray", c["nodes"][k] / r, "tris/ray", c["tris"][k] / r, "B_ray", 32 * c["nodes"][k] / r + 48 * c["tris"][k] / r + 64)
