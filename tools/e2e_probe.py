import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32
desc = scenes.c2_spot(); bvh = Bvh.Build(desc.prims)
tex = np.zeros((desc.width, desc.height, 4))
for k in range(4):
    t0 = time.perf_counter(); sc = Scene(desc, bvh=bvh); t1 = time.perf_counter()
    integ = CudaPixelIntegrator(sc, precision=FAST_F32, seed=1)
    integ.Sample(64, out=tex); t2 = time.perf_counter()
    st = integ.stats
    sc.close(); t3 = time.perf_counter()
    print(f"iter {k}: Scene() {1e3*(t1-t0):.1f} ms, Sample->host {1e3*(t2-t1):.1f} ms (device {st['ms_total']:.1f} ms), close {1e3*(t3-t2):.1f} ms", flush=True)
