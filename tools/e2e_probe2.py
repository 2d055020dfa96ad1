#!/usr/bin/env python3
"""Host-timed C2 frames (1920x1080, 64 spp -> pinned Color[w,h]) four ways: scene kept or re-created per frame, blocking
or pipelined (two frames in flight).  Prints ms per frame next to the device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32, _lib
desc = scenes.c2_spot(); bvh = Bvh.Build(desc.prims)
spp, K = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), (int(sys.argv[2]) if len(sys.argv) > 2 else 10)
lib = _lib.load()
texs = [np.zeros((desc.width, desc.height, 4)) for _ in range(2)]
for t in texs:
    _lib.check(lib.mfx_host_register(_lib.ptr(t), t.nbytes))
sc = Scene(desc, bvh=bvh); integ = CudaPixelIntegrator(sc, precision=FAST_F32, seed=1)
for _ in range(3):
    integ.Sample(spp, out=texs[0])
dev = integ.stats["ms_total"]
t0 = time.perf_counter()
for k in range(K):
    integ.Sample(spp, out=texs[0])
a = (time.perf_counter() - t0) / K * 1e3
integ.SampleAsync(spp, texs[0]); integ.SampleAsync(spp, texs[1]); integ.Wait(); integ.Wait()
t0 = time.perf_counter()
integ.SampleAsync(spp, texs[0])
for k in range(1, K):
    integ.SampleAsync(spp, texs[k % 2]); integ.Wait()
integ.Wait()
b = (time.perf_counter() - t0) / K * 1e3
sc.close()
def recreated(pipe, n):
    prev = None
    for k in range(n):
        s2 = Scene(desc, bvh=bvh); i2 = CudaPixelIntegrator(s2, precision=FAST_F32, seed=1)
        if pipe:
            i2.SampleAsync(spp, texs[k % 2])
            if prev: prev[1].Wait(); prev[0].close()
            prev = (s2, i2)
        else:
            i2.Sample(spp, out=texs[0]); s2.close()
    if prev: prev[1].Wait(); prev[0].close()
recreated(True, 3); recreated(False, 2)
t0 = time.perf_counter(); recreated(False, K); c = (time.perf_counter() - t0) / K * 1e3
t0 = time.perf_counter(); recreated(True, K); d = (time.perf_counter() - t0) / K * 1e3
print(f"device {dev:.2f} ms | kept scene: blocking {a:.2f}, pipelined {b:.2f} | re-created scene: blocking {c:.2f}, pipelined {d:.2f}")
