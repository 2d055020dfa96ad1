import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
v = sys.argv[1]
os.environ["MFX_TRACE_VARIANT"] = v
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, EXACT_F64
for name, kw, spp in [("cornell", {}, 4), ("c1_cube", dict(width=320, height=240), 4), ("c2_spot", dict(width=480, height=270), 4), ("c3_renault", dict(width=480, height=270), 4), ("c4_spheres", dict(width=480, height=270, grid=60), 4)]:
    desc = scenes.WORKLOADS[name](**kw)
    s = Scene(desc)
    f = CudaPixelIntegrator(s, precision=FAST_F32, seed=1).Sample(spp).copy()
    e = CudaPixelIntegrator(s, precision=EXACT_F64, seed=1).Sample(spp).copy()
    clip = np.percentile(np.abs(e[:, :, :3]), 99.5)
    c = lambda x: np.clip(x[:, :, :3], -clip, clip)
    rel = np.sqrt(((c(f) - c(e)) ** 2).mean()) / np.abs(c(e)).mean()
    print(f"variant {v} {name:11s} rel_rmse(clipped)={rel:.3e} mean_ratio={c(f).mean()/c(e).mean():.5f}", flush=True)
