#!/usr/bin/env python3
"""A/B of the bounce-0 kernel: id-exact hybrid (default) against f32 primitive tests (MFX_SAMPLE_F32_PRIMARY), on the
frame (all bounces) and on primary rays alone (max_depth 0).  usage: hyb_probe.py [workload,...] [spp]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, _lib

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2_spot", "c3_renault"]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for name in names:
    for depth0 in (True, False):
        desc = scenes.WORKLOADS[name]()
        if depth0:
            desc.max_depth = 0
        s = Scene(desc)
        integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
        row = {"workload": name, "max_depth": desc.max_depth, "spp": spp}
        for label, flags in (("hybrid", 0), ("f32", _lib.SAMPLE_F32_PRIMARY)):
            best = None
            for _ in range(3):
                integ.SampleF32(spp, flags=flags)
                st = integ.stats
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = dict(st)
            rays = best["closest_rays"] + best["shadow_rays"]
            row[label] = {"ms_total": round(best["ms_total"], 3), "ms_extend": round(best["ms_extend"], 3), "mrays_s": round(rays / best["ms_total"] / 1e3, 1),
                          "extend_mrays_s": round(best["closest_rays"] / best["ms_extend"] / 1e3, 1), "fixups": best["hybrid_fixups"]}
        print(json.dumps(row), flush=True)
        s.close()
