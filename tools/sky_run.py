#!/usr/bin/env python3
"""One fast-path frame of the sphere sample (MFX_SKY_TRACER, RandomScene at 1920x1080) -- the command the ncu
captures of k_f_shade_sky / k_f_trace6 in profiles/ were taken from.  usage: sky_run.py [spp] [repeats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = Scene(scenes.random_scene(width=1920, height=1080))
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
for _ in range(rep):
    integ.SampleF32(spp)
st = integ.stats
print(f"random_scene 1920x1080 x {spp} spp: {st['closest_rays']} rays, {st['ms_total']:.2f} ms, "
      f"{st['closest_rays'] / st['ms_total'] / 1e3:.0f} Mrays/s, extend {st['closest_rays'] / st['ms_extend'] / 1e3:.0f} Mrays/s, {st['launches']} launches")
