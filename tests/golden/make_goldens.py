#!/usr/bin/env python3
"""Mint tests/golden/oracle_goldens.npz from the CPU oracle (oracle/mafrix_oracle.c).

The reference ships no golden vectors, so these pin the ORACLE against drift and
give the -m gpu tests committed vectors to compare with besides the live oracle.  They are
outputs of the restatement, not of the F# program: parity stays "unpinned" (DESIGN.md).

Run:  python tests/golden/make_goldens.py        (everything)
      python tests/golden/make_goldens.py sky    (only tests/golden/sky_goldens.npz, the sphere sample)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mafrixraytracing_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402

PRIMARY = {   # name -> builder kwargs (pixel-centre primary-hit buffers)
    "cornell": {}, "c1_cube": {}, "c2_spot_small": dict(width=480, height=270),
    "c3_renault_small": dict(width=480, height=270), "c2_spot": {}, "c3_renault": {},
}
IMAGES = {    # name -> (builder kwargs, spp, seed)
    "cornell": (dict(width=100, height=100), 2, 7),
    "c1_cube": (dict(width=80, height=60), 4, 7),
    "c2_spot": (dict(width=96, height=54), 2, 7),
    "c3_renault": (dict(width=96, height=54), 2, 7),
    "spheres": (dict(width=96, height=54, grid=24), 2, 7),
}


def builder(name):
    base = name.replace("_small", "")
    return scenes.c4_spheres if base == "spheres" else scenes.WORKLOADS[base]


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    out = {}
    for name, kw in PRIMARY.items():
        o = oracle.OracleScene(builder(name)(**kw))
        prim, t = o.trace_primary()
        out[f"primary/{name}/sha_prim"] = sha(prim)
        out[f"primary/{name}/sha_t"] = sha(t)
        out[f"primary/{name}/hits"] = np.int64((prim >= 0).sum())
        if prim.size <= 320000:
            out[f"primary/{name}/prim"] = prim
            out[f"primary/{name}/t_stride97"] = t[::97]
        print(name, prim.size, int((prim >= 0).sum()))
    for name, (kw, spp, seed) in IMAGES.items():
        o = oracle.OracleScene(builder(name)(**kw))
        tex = o.sample(spp, seed=seed)
        out[f"image/{name}/rgb"] = tex[:, :, :3].copy()
        out[f"image/{name}/spp_seed"] = np.array([spp, seed])
        print("image", name, tex.shape, tex[:, :, :3].mean())
    out["philox/zero"] = oracle.philox([0, 0, 0, 0], [0, 0])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_goldens.npz"), **out)


def sky_scenes():
    """The sphere sample's scenes (MFX_SKY_TRACER): RandomScene as shipped, and a small scene with every material,
    both textures and a finite aperture."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.test_oracle_sky import MIXED, sky_desc
    from mafrixraytracing_b200.scene import RayTraceCamera
    rf, pm = scenes.perlin_tables(3)
    cam = RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, 1.5, 0.4, 3.4)
    return {"random_scene": scenes.random_scene(),
            "mixed": sky_desc(MIXED["centers"], MIXED["radii"], MIXED["specs"], width=96, height=64, cam=cam, tables=(rf, pm))}


def main_sky():
    """tests/golden/sky_goldens.npz: outputs of oracle/mafrix_oracle_sky.c (same caveat: they pin the restatement)."""
    out = {}
    for name, desc in sky_scenes().items():
        o = oracle.OracleSkyScene(desc)
        prim, t = o.trace_primary()
        out[f"primary/{name}/prim"] = prim
        out[f"primary/{name}/sha_t"] = sha(t)
        out[f"primary/{name}/t_stride97"] = t[::97]
        if name == "random_scene":                         # image golden at 120x60: same spheres, smaller film
            o = oracle.OracleSkyScene(scenes.random_scene(width=120, height=60))
        tex = o.sample(2, seed=7)
        out[f"image/{name}/rgb"] = tex[:, :, :3].copy()
        out[f"image/{name}/display_sha"] = sha(oracle.sky_display_rgba8(tex))
        print("sky", name, prim.size, int((prim >= 0).sum()), tex[:, :, :3].mean())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sky_goldens.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] != "sky":       # `make_goldens.py sky` mints only the sky file
        main()
    main_sky()
