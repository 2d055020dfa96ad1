import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, EXACT_F64, FAST_F32
def desc_of(name, kw):
    return (scenes.c4_spheres if name == "spheres" else scenes.WORKLOADS[name])(**kw)
for name, kw in (("c3_renault", dict(width=160, height=90)), ("spheres", dict(width=160, height=90, grid=24)), ("c3_renault", dict(width=480, height=270))):
    s = Scene(desc_of(name, kw))
    for spp in (64, 256):
        ea = CudaPixelIntegrator(s, precision=EXACT_F64, seed=5).Sample(spp).copy()[:, :, :3]
        eb = CudaPixelIntegrator(s, precision=EXACT_F64, seed=6).Sample(spp).copy()[:, :, :3]
        fa = CudaPixelIntegrator(s, precision=FAST_F32, seed=5).Sample(spp).copy()[:, :, :3]
        clip = np.percentile(ea, 99.0)
        c = lambda x: np.clip(x, -clip, clip)
        lum = lambda x: (c(x) * [0.2126, 0.7152, 0.0722]).sum(-1)
        noise = np.sqrt(((c(ea) - c(eb)) ** 2).mean()); err = np.sqrt(((c(fa) - c(ea)) ** 2).mean())
        m = np.abs(c(ea)).mean()
        print(name, kw, spp, f"rel_rmse fast-vs-exact {err/m:.4f}  exact-vs-exact {noise/m:.4f}  mean ratio {c(fa).mean()/c(ea).mean():.5f}  rel luminance diff {abs(lum(fa).mean()/lum(ea).mean()-1):.5f}", flush=True)
