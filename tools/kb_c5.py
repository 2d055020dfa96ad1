import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32
import mafrixraytracing_b200._lib as L
desc = scenes.c5_soup(); bvh = Bvh.Build(desc.prims)
for v in sys.argv[1].split(","):
    os.environ["MFX_TRACE_VARIANT"] = v
    s = Scene(desc, bvh=bvh)
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    for rep in range(2):
        integ.SampleF32(1)
    st = integ.stats
    rays = st["closest_rays"] + st["shadow_rays"]
    print(f"variant {v}: C5 {rays/st['ms_total']/1e3:.1f} Mrays/s (extend {st['closest_rays']/st['ms_extend']/1e3:.1f}, shadow {st['shadow_rays']/st['ms_shadow']/1e3:.1f})", flush=True)
    s.close()
