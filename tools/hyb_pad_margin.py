#!/usr/bin/env python3
"""How much margin does the per-ray box pad of the id-exact kernel have?  Sweeps MFX_HYB_PAD_PPB (default 4000 = 4e-6)
down to zero on scenes chosen to hurt (far from the origin, vertex-aimed rays) and counts closest hits that differ from
the exact kernel's.  The shipped factor must sit well above the first value that shows a mismatch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, FAST_F32, EXACT_F64
from tests.test_gpu_hybrid_stress import _soup
from mafrixraytracing_b200.scene import AreaLight, PinholeCamera, SceneDesc, make_materials

def cases():
    yield "c2_spot 1080p pixel centres", scenes.c2_spot(), None
    yield "c3_renault 1080p pixel centres", scenes.c3_renault(), None
    for seed, off, scale, grid in ((5, (5000.0, -3000.0, 8000.0), 1.0, False), (6, (100.0, 100.0, 100.0), 0.01, True), (2, (0, 0, 0), 1.0, True)):
        rng = np.random.default_rng(seed)
        off = np.array(off, float)
        prims = _soup(rng, 3000, 0, 0, scale, off, grid)
        light = AreaLight(np.array([(-1, 3, 1), (-1, 3, -1), (1, 3, -1), (1, 3, 1)], float) * scale + off, (0, -1, 0), (10, 10, 10))
        cam = PinholeCamera(off + np.array([0.3, 0.8, 3.5]) * scale, (-0.05, -0.2, -1), 120.0, 1.0)
        yield f"soup seed {seed} offset {tuple(off)} scale {scale}", SceneDesc(prims, make_materials([("lambert", (0.7,) * 3)]), light, cam, 64, 64, 2, 0), rng.random((400000, 2))

for name, desc, uv in cases():
    row = {"scene": name}
    base = Scene(desc)
    ep, et = base.TracePrimary(uv, precision=EXACT_F64)
    base.close()
    for ppb in (4000, 400, 40, 4, 0):
        os.environ["MFX_HYB_PAD_PPB"] = str(ppb)
        s = Scene(desc)
        fp, ft = s.TracePrimary(uv, precision=FAST_F32)
        row[f"pad_{ppb}ppb_mismatches"] = int(((fp != ep) | (ft != et)).sum())
        s.close()
    row["rays"] = len(ep)
    print(json.dumps(row), flush=True)
