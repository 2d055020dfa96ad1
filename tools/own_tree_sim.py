#!/usr/bin/env python3
"""Writes the triangles of a workload (default C2: spot + floor + back wall, Rects as two triangles) as n x 9 doubles
for tools/own_tree_sim.cpp, builds the simulator and runs it.  CPU only.  usage: own_tree_sim.py [workload] [w h]"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mafrixraytracing_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "c2_spot"
desc = scenes.WORKLOADS[name]()
tris = []
for q in desc.prims:
    v = q["v"]
    if q["kind"] == 2:
        raise SystemExit("triangles and rects only")
    tris.append(v[:9])
    if q["kind"] == 1:
        tris.append(np.concatenate([v[0:3], v[6:9], v[9:12]]))
t = np.array(tris, np.float64)
td = tempfile.mkdtemp()
t.tofile(os.path.join(td, "tris.bin"))
exe = os.path.join(td, "own_tree_sim")
csrc = os.path.join(ROOT, "mafrixraytracing_b200", "csrc")
subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-I", "/usr/local/cuda/include", "-I", csrc, "-o", exe,
                       os.path.join(ROOT, "tools", "own_tree_sim.cpp"), os.path.join(csrc, "mfx_build.cpp")])
wh = sys.argv[2:4] if len(sys.argv) > 3 else ["480", "270"]
cam, lp = desc.camera, desc.light.p
view = [*cam.pos_arg, *cam.dir_arg, cam.fov, cam.aspect, *lp[0], *lp[1], *lp[3], min(desc.max_depth, 5)]
subprocess.check_call([exe, os.path.join(td, "tris.bin"), str(len(t))] + wh + [repr(float(x)) for x in view])
