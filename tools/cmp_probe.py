#!/usr/bin/env python3
"""64-byte quantised records (QuadC, MFX_TRACE_VARIANT=7) against the shipped 128-byte records: records and primitive
tests per ray from the instrumented kernels (MFX_SAMPLE_COUNT_OWN_TREE), frame times, pixels that differ."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, _lib
for name, spp in (("c2_spot", 16), ("c3_renault", 16), ("c4_spheres", 4), ("c5_soup", 2)):
    if name not in scenes.WORKLOADS:
        continue
    s = Scene(scenes.WORKLOADS[name]())
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    row, ref = {"workload": name, "spp": spp}, None
    for v in ("0", "7"):
        os.environ["MFX_TRACE_VARIANT"] = v
        best = None
        for _ in range(3):
            img = integ.SampleF32(spp)
            st = integ.stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best = dict(st)
        integ.Sample(1, flags=_lib.SAMPLE_COUNT_OWN_TREE)
        c = integ.stats
        rays = best["closest_rays"] + best["shadow_rays"]
        row["records_128B" if v == "0" else "records_64B"] = {
            "mrays_s": round(rays / best["ms_total"] / 1e3, 1), "ms_extend": round(best["ms_extend"], 2), "ms_shadow": round(best["ms_shadow"], 2),
            "closest_records_per_ray": round(c["nodes"][0] / max(1, c["closest_rays"]), 3), "closest_prim_tests_per_ray": round((c["tris"][0] + c["spheres"][0]) / max(1, c["closest_rays"]), 3),
            "shadow_records_per_ray": round(c["nodes"][1] / max(1, c["shadow_rays"]), 3), "shadow_prim_tests_per_ray": round((c["tris"][1] + c["spheres"][1]) / max(1, c["shadow_rays"]), 3)}
        if ref is None:
            ref = img.copy()
        else:
            row["pixels_differing"] = int((np.abs(ref - img).max(axis=-1) > 0).sum())
    print(json.dumps(row), flush=True)
    s.close()
