#!/usr/bin/env python3
"""The C5 workload (spot x 1708 = 10.0 M triangles, 3840x2160, the one config whose scene -- 1.1 GB -- leaves L2) for ncu:
two warm 2-spp frames, then one 2-spp frame whose launches are the ones to read.  Prints the own-tree record / triangle
counts per ray (the requested bytes of the HBM roofline) and the device times of the last frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, FAST_F32, _lib
desc = scenes.c5_soup(); bvh = Bvh.Build(desc.prims)
s = Scene(desc, bvh=bvh)
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(3):
    integ.SampleF32(spp)
st = dict(integ.stats)
integ.SampleF32(spp, flags=_lib.SAMPLE_COUNT_OWN_TREE)
c = integ.stats
out = {"spp": spp, "closest_rays": st["closest_rays"], "shadow_rays": st["shadow_rays"], "ms_total": st["ms_total"], "ms_extend": st["ms_extend"],
       "ms_shadow": st["ms_shadow"], "launches": st["launches"],
       "records_per_ray": [c["nodes"][0] / c["closest_rays"], c["nodes"][1] / c["shadow_rays"]],
       "tris_per_ray": [c["tris"][0] / c["closest_rays"], c["tris"][1] / c["shadow_rays"]], "scene_bytes": s.device_bytes()}
print(json.dumps(out))
