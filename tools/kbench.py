#!/usr/bin/env python3
"""Kernel A/B harness: times the fast path on a workload for several MFX_TRACE_VARIANT values
and checks that every variant produces the same frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def run(variant, name, spp, kw, reps=3):
    os.environ["MFX_TRACE_VARIANT"] = str(variant)
    from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32
    desc = scenes.WORKLOADS[name](**kw)
    s = Scene(desc)
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    best = None
    for _ in range(reps):
        img = integ.SampleF32(spp)
        st = integ.stats
        if best is None or st["ms_total"] < best["ms_total"]:
            best = dict(st)
    rays = best["closest_rays"] + best["shadow_rays"]
    print(f"variant={variant:2d} {name} spp={spp} Mrays/s={rays/best['ms_total']/1e3:8.1f} total_ms={best['ms_total']:7.2f} "
          f"extend_ms={best['ms_extend']:7.2f} ({best['closest_rays']/best['ms_extend']/1e3:7.1f} M/s) shadow_ms={best['ms_shadow']:7.2f} "
          f"({best['shadow_rays']/best['ms_shadow']/1e3:7.1f} M/s) other_ms={best['ms_shade']:6.2f}", flush=True)
    s.close()
    return img

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_spot"
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 5, 6, 7, -1]
    ref = None
    for v in variants:
        img = run(v, name, spp, {})
        if ref is None:
            ref = img
        else:
            same = np.array_equal(img, ref)
            if not same:
                d = np.abs(img - ref)
                print(f"   variant {v}: frame differs from variant {variants[0]}: max_abs={d.max():.3e} n={int((d>0).any(axis=2).sum())}")
