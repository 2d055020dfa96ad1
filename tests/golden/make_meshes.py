#!/usr/bin/env python3
"""Mint the mesh fixtures used by the tests and the bench.

/root/reference does not exist on the GPU box, so the three OBJ models the
BASELINE configs name (3DModel/Cube, 3DModel/spot, 3DModel/Renault12TL) are
converted ONCE, in the build container, into compact binary fixtures:

    tests/golden/meshes/<name>.npz   v : float64 [nv,3]   (text -> double, as .NET `float` parses)
                                     f : int32   [nf,4]   (0-based; 4th index -1 for a triangle)

Face semantics follow the reference loader (EngineCore/Models/ObjModelLoader.fs:63-92):
3 vertices -> Triangle, 4 -> Rect, index i>0 -> i-1, i<0 -> len+i; only the
geometric-vertex index of `a`, `a/b`, `a//c`, `a/b/c` is used.

Run:  python tests/golden/make_meshes.py  (needs /root/reference)
"""
import os, sys
import numpy as np

REF = "/root/reference/3DModel"
MODELS = {
    "cube": "Cube/Cube.obj",
    "spot": "spot/spot_triangulated_good.obj",
    "renault": "Renault12TL/Renault12TL.obj",
}

def parse_obj(path):
    verts, faces = [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == "f":
                idx = []
                for ref in tok[1:]:
                    i = int(ref.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                assert len(idx) in (3, 4), f"{path}: face with {len(idx)} vertices"
                faces.append(idx + [-1] * (4 - len(idx)))
    return np.asarray(verts, np.float64), np.asarray(faces, np.int32)

def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes")
    os.makedirs(out, exist_ok=True)
    for name, rel in MODELS.items():
        v, f = parse_obj(os.path.join(REF, rel))
        np.savez_compressed(os.path.join(out, name + ".npz"), v=v, f=f)
        print(f"{name}: {len(v)} vertices, {len(f)} faces "
              f"({int((f[:,3] < 0).sum())} tri, {int((f[:,3] >= 0).sum())} quad), "
              f"extent {v.min(0)} .. {v.max(0)}")

if __name__ == "__main__":
    sys.exit(main())
