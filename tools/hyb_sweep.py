#!/usr/bin/env python3
"""Sweeps MFX_HYB_VARIANT (vote threshold / blocks per SM of the id-exact primary kernel) on primary rays alone."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, _lib
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2_spot", "c3_renault"]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for name in names:
    desc = scenes.WORKLOADS[name]()
    desc.max_depth = 0
    s = Scene(desc)
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    for v in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]:
        os.environ["MFX_HYB_VARIANT"] = str(v)
        best = None
        for _ in range(3):
            integ.SampleF32(spp)
            st = integ.stats
            if best is None or st["ms_extend"] < best["ms_extend"]:
                best = dict(st)
        print(json.dumps({"workload": name, "variant": v, "ms_total": round(best["ms_total"], 3), "ms_extend": round(best["ms_extend"], 3),
                          "primary_mrays_s": round(best["closest_rays"] / best["ms_extend"] / 1e3, 1)}), flush=True)
    s.close()
