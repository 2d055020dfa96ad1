#!/usr/bin/env python3
"""Host-side anatomy of SMALL frames (what one of eight GPUs renders: 1/8 of the C2 frame's paths): device time against the
blocking call on a kept scene.  usage: small_frame_probe.py [spp]   (MFX_DEBUG=2 prints the library's own timestamps)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32, _lib
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
desc = scenes.c2_spot(); bvh = Bvh.Build(desc.prims)
tex = np.zeros((desc.width, desc.height, 4))
_lib.check(_lib.load().mfx_host_register(_lib.ptr(tex), tex.nbytes))
sc = Scene(desc, bvh=bvh); integ = CudaPixelIntegrator(sc, precision=FAST_F32, seed=1)
frame = None
for k in range(6):
    t0 = time.perf_counter()
    integ.SampleF32(spp)
    t1 = time.perf_counter()
    print(f"SampleF32({spp}): host {1e3 * (t1 - t0):.3f} ms, device {integ.stats['ms_total']:.3f} ms, launches {integ.stats['launches']}", flush=True)
