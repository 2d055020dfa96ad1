#!/usr/bin/env python3
"""The sphere sample's late bounces: one launch pair per bounce to depth 50 (MFX_SKY_TAIL=0) against the tail launch
(default: from bounce 8 on every surviving path is traced and shaded to its end by k_f_trace6<TAIL>).  Frames must be
identical; prints ms, Mrays/s and launches at the sample's own size and at 1080p."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32
for (w, h, spp) in ((400, 200, 9), (400, 200, 100), (1920, 1080, 32)):
    s = Scene(scenes.random_scene(width=w, height=h))
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    row, ref = {"frame": f"{w}x{h}", "spp": spp}, None
    for tail in ("0", "8", "4", "12"):
        os.environ["MFX_SKY_TAIL"] = tail
        best, host = None, None
        for _ in range(4):
            t0 = time.perf_counter()
            img = integ.SampleF32(spp)
            dt = (time.perf_counter() - t0) * 1e3
            st = integ.stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best, host = dict(st), dt
        row["tail_from_" + tail if tail != "0" else "per_bounce"] = {"ms": round(best["ms_total"], 3), "mrays_s": round(best["closest_rays"] / best["ms_total"] / 1e3, 1), "launches": best["launches"]}
        if ref is None:
            ref = img.copy()
        else:
            row["identical"] = row.get("identical", True) and bool(np.array_equal(ref, img))
    print(json.dumps(row), flush=True)
    s.close()
