#!/usr/bin/env python3
"""One-off validation + timing of C5 (spot instanced 1708x = 10.0M triangles, 3840x2160): primary-hit
parity of both precisions against the oracle on 200k jittered rays, then fast-path timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, EXACT_F64, FAST_F32
from oracle import oracle
t0 = time.time(); desc = scenes.c5_soup(); print(f"scene: {len(desc.prims)} prims in {time.time()-t0:.1f}s", flush=True)
t0 = time.time(); bvh = Bvh.Build(desc.prims); print(f"mfx_bvh_build: {time.time()-t0:.1f}s", flush=True)
t0 = time.time(); s = Scene(desc, bvh=bvh); print(f"scene_create: {time.time()-t0:.1f}s", flush=True)
rng = np.random.default_rng(1)
uv = rng.random((200000, 2))
t0 = time.time(); o = oracle.OracleScene(desc); print(f"oracle build: {time.time()-t0:.1f}s", flush=True)
nodes, idx = o.bvh()
print("trees equal:", bool(np.array_equal(idx, bvh.indices) and np.array_equal(nodes.view(np.uint8), bvh.nodes.view(np.uint8))), flush=True)
t0 = time.time(); op, ot = o.trace_primary(uv); print(f"oracle primary 200k: {time.time()-t0:.1f}s hits={(op>=0).sum()}", flush=True)
ep, et = s.TracePrimary(uv, precision=EXACT_F64)
print("exact: id_mismatch", int((ep != op).sum()), "t_exact", bool(np.array_equal(et, ot)), flush=True)
fp, ft = s.TracePrimary(uv, precision=FAST_F32)
both = (fp == op) & (op >= 0)
print("fast: id_mismatch", int((fp != op).sum()), "max_rel_t", float((np.abs(ft[both]-ot[both])/ot[both]).max()), flush=True)
print("device bytes", s.device_bytes(), flush=True)
integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
for rep in range(2):
    integ.SampleF32(1)
    st = integ.stats
    rays = st["closest_rays"] + st["shadow_rays"]
    print(f"C5 4K x1spp: {rays/1e6:.1f}M rays {st['ms_total']:.1f} ms -> {rays/st['ms_total']/1e3:.1f} Mrays/s (extend {st['closest_rays']/st['ms_extend']/1e3:.1f}, shadow {st['shadow_rays']/st['ms_shadow']/1e3:.1f})", flush=True)
