#!/usr/bin/env python3
"""MFX_EXACT_F64 frames: closest hits through the hybrid kernel (default) against the plain walk of the reference tree
(MFX_EXACT_WALK=1).  Frames must be identical; prints Mrays/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, EXACT_F64
for name, kw, spp in (("c2_spot", {}, 2), ("c3_renault", {}, 2), ("c4_spheres", {}, 1), ("cornell", {}, 16), ("random_scene", dict(width=1920, height=1080), 2)):
    desc = scenes.WORKLOADS[name](**kw)
    s = Scene(desc)
    integ = CudaPixelIntegrator(s, precision=EXACT_F64, seed=1)
    row, ref = {"workload": name, "spp": spp}, None
    for walk in ("1", "0"):
        os.environ["MFX_EXACT_WALK"] = walk
        best = None
        for _ in range(2):
            img = integ.Sample(spp).copy()
            st = integ.stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best = dict(st)
        rays = best["closest_rays"] + best["shadow_rays"]
        row["reference_walk" if walk == "1" else "hybrid_closest"] = {"mrays_s": round(rays / best["ms_total"] / 1e3, 1), "ms_extend": round(best["ms_extend"], 2),
                                                                      "ms_shadow": round(best["ms_shadow"], 2), "ms_total": round(best["ms_total"], 2), "fixups": best["hybrid_fixups"]}
        if ref is None:
            ref = img
        else:
            row["identical_frames"] = bool(np.array_equal(ref, img))
    print(json.dumps(row), flush=True)
    s.close()
