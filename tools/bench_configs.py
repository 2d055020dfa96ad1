#!/usr/bin/env python3
"""Runs every BASELINE config (C1..C5) through the fast and the exact path once and writes a table
(profiles/<tag>_configs.json): rays, device ms, Mrays/s, per-class Mrays/s, scene bytes.  spp is
reduced for the big configs (cost is linear in spp); the headline bench stays bench.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, EXACT_F64, FAST_F32

RUNS = [("cornell", 256, 16), ("c1_cube", 64, 16), ("c2_spot", 64, 2), ("c3_renault", 64, 2), ("c4_spheres", 8, 1), ("c5_soup", 8, 1),
        ("random_scene", 4096, 64), ("random_scene_1080p", 64, 2)]      # the sphere sample (MFX_SKY_TRACER), as shipped and at 1080p
out = []
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name, spp_fast, spp_exact in RUNS:
    if only and name not in only:
        continue
    t0 = time.time()
    desc = scenes.random_scene(width=1920, height=1080) if name == "random_scene_1080p" else scenes.WORKLOADS[name]()
    t_scene = time.time() - t0
    t0 = time.time(); bvh = Bvh.Build(desc.prims); t_bvh = time.time() - t0
    s = Scene(desc, bvh=bvh)
    row = {"config": name, "prims": int(len(desc.prims)), "size": [desc.width, desc.height], "max_depth": desc.max_depth,
           "integrator": ["PathIntegrator", "NewPathTracer", "GetColor (sky tracer)"][desc.integrator], "host_bvh_build_s": round(t_bvh, 2)}
    for prec, label, spp in ((FAST_F32, "fast_f32", spp_fast), (EXACT_F64, "exact_f64", spp_exact)):
        if name == "c5_soup" and prec == EXACT_F64:
            continue
        integ = CudaPixelIntegrator(s, precision=prec, seed=1)
        t0 = time.time(); integ.SampleF32(1); t_first = time.time() - t0      # includes the layout build (fast: own SAH tree)
        best = None
        for _ in range(2):
            integ.SampleF32(spp)
            st = integ.stats
            if best is None or st["ms_total"] < best["ms_total"]:
                best = dict(st)
        rays = best["closest_rays"] + best["shadow_rays"]
        row[label] = {"spp": spp, "rays": rays, "ms": round(best["ms_total"], 3), "mrays_s": round(rays / best["ms_total"] / 1e3, 1),
                      "extend_mrays_s": round(best["closest_rays"] / max(best["ms_extend"], 1e-9) / 1e3, 1),
                      "shadow_mrays_s": round(best["shadow_rays"] / max(best["ms_shadow"], 1e-9) / 1e3, 1),
                      "spp_per_s": round(spp / (best["ms_total"] * 1e-3), 2), "first_call_s": round(t_first, 2)}
    row["device_bytes"] = s.device_bytes() if name != "c5_soup" else None
    print(json.dumps(row), flush=True)
    out.append(row)
    s.close()
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/{tag}_configs.json", "w"), indent=1)
