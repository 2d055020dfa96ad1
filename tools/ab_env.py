#!/usr/bin/env python3
"""A/B of an environment knob on whole frames.  usage: ab_env.py VAR v1,v2,.. workload spp[,spp..] [reps] [rebuild]
rebuild: the knob is read when the scene's layouts are built (tree builder knobs): a fresh Scene per value, tree cache off."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32
var, vals, name = sys.argv[1], sys.argv[2].split(","), sys.argv[3]
spps = [int(x) for x in sys.argv[4].split(",")]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
rebuild = len(sys.argv) > 6
if rebuild:
    os.environ["MFX_TREE_CACHE"] = "0"
desc = scenes.WORKLOADS[name]()
s = Scene(desc)
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
ref = {}
for spp in spps:
    for v in vals:
        os.environ[var] = v
        if rebuild:
            s.close()
            s = Scene(desc)
            integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
        best = None
        for _ in range(reps):
            img = integ.SampleF32(spp)
            st = integ.stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best = dict(st)
        same, ndiff = None, None
        if spp in ref:
            same = bool(np.array_equal(ref[spp], img))
            ndiff = int((np.abs(ref[spp] - img).max(axis=-1) > 0).sum())
        else:
            ref[spp] = img.copy()
        rays = best["closest_rays"] + best["shadow_rays"]
        print(json.dumps({"workload": name, "spp": spp, var: v, "ms_total": round(best["ms_total"], 3), "mrays_s": round(rays / best["ms_total"] / 1e3, 1),
                          "ms_extend": round(best["ms_extend"], 3), "ms_shadow": round(best["ms_shadow"], 3), "same_frame_as_first": same, "pixels_differing": ndiff}), flush=True)
