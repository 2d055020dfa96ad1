#!/usr/bin/env python3
"""Where does a mfx_multi_create + mfx_multi_sample + destroy step spend its host time?  usage: multi_probe.py [n_gpus] [spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Bvh, MultiGpuPixelIntegrator, FAST_F32, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else _lib.load().mfx_device_count()
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
desc = scenes.c2_spot(); bvh = Bvh.Build(desc.prims)
tex = np.zeros((desc.width, desc.height, 4))
_lib.check(_lib.load().mfx_host_register(_lib.ptr(tex), tex.nbytes))
for k in range(4):
    t0 = time.perf_counter()
    m = MultiGpuPixelIntegrator(desc, devices=list(range(n)), bvh=bvh, precision=FAST_F32, seed=1)
    t1 = time.perf_counter()
    m.Sample(spp, out=tex)
    t2 = time.perf_counter()
    m.Sample(spp, out=tex)
    t3 = time.perf_counter()
    dev = max(p["ms_total"] for p in m.stats["per_device"])
    m.close()
    t4 = time.perf_counter()
    print(f"iter {k}: create {1e3*(t1-t0):.2f} ms, first sample {1e3*(t2-t1):.2f} ms, second sample {1e3*(t3-t2):.2f} ms (slowest device {dev:.2f} ms), destroy {1e3*(t4-t3):.2f} ms", flush=True)

# pipelined: frame k renders while the host creates the replicas of frame k+1 -- where does the host thread's time go?
texs = [tex, np.zeros_like(tex)]
_lib.check(_lib.load().mfx_host_register(_lib.ptr(texs[1]), texs[1].nbytes))
acc = {"create": 0.0, "post": 0.0, "wait": 0.0, "close": 0.0}
K = 30
stamps = []
prev = None
t_all = t_all0 = time.perf_counter()
for k in range(K):
    t0 = time.perf_counter()
    m = MultiGpuPixelIntegrator(desc, devices=list(range(n)), bvh=bvh, precision=FAST_F32, seed=1)
    t1 = time.perf_counter()
    m.SampleAsync(spp, texs[k % 2])
    t2 = time.perf_counter()
    if prev is not None:
        prev.Wait()
        t3 = time.perf_counter()
        prev.close()
        t4 = time.perf_counter()
        acc["wait"] += t3 - t2; acc["close"] += t4 - t3
        stamps.append(t3)
    acc["create"] += t1 - t0; acc["post"] += t2 - t1
    prev = m
prev.Wait(); prev.close()
t_all = time.perf_counter() - t_all
print("frame completion intervals, ms:", " ".join(f"{1e3 * (b - a):.1f}" for a, b in zip([t_all0] + stamps[:-1], stamps)))
steady = (stamps[-1] - stamps[4]) / (len(stamps) - 5)
print(f"pipelined, steady state (frames 5..{K - 1}): {1e3 * steady:.2f} ms per frame")
print(f"pipelined x{K}: {1e3 * t_all / K:.2f} ms per frame; host thread per frame: " + ", ".join(f"{k} {1e3 * v / K:.2f}" for k, v in acc.items()), flush=True)
