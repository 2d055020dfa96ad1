#!/usr/bin/env python3
"""Mint tests/golden/big_goldens.npz: pixel-centre primary-hit buffers of BASELINE configs[3] / [4] AT THEIR STATED SIZE
(C4: 99 857 spheres, 3840x2160; C5: spot x 1708 = 10.0 M triangles, 3840x2160) from the CPU oracle -- checksums plus a
strided sample for diagnostics -- and the primary buffer of C3 at 1920x1080 under its own integrator (already in
oracle_goldens.npz as primary/c3_renault).  Same caveat as make_goldens.py: outputs of the restatement.

Run:  python tests/golden/make_goldens_big.py      (C5 needs ~8 GB and a few minutes on 8 cores)"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mafrixraytracing_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    out = {}
    for name in ("c4_spheres", "c5_soup"):
        t0 = time.time()
        desc = scenes.WORKLOADS[name]()
        o = oracle.OracleScene(desc)
        prim, t = o.trace_primary()
        out[f"primary/{name}/sha_prim"] = sha(prim)
        out[f"primary/{name}/sha_t"] = sha(t)
        out[f"primary/{name}/hits"] = np.int64((prim >= 0).sum())
        out[f"primary/{name}/prim_stride997"] = prim[::997].copy()
        out[f"primary/{name}/t_stride997"] = t[::997].copy()
        # 200 k jittered rays (seed 1): checksums of the oracle's answer, so the GPU test needs no CPU oracle run at this size
        uv = np.random.default_rng(1).random((200000, 2))
        jp, jt = o.trace_primary(uv)
        out[f"jitter/{name}/sha_prim"] = sha(jp)
        out[f"jitter/{name}/sha_t"] = sha(jt)
        out[f"jitter/{name}/hits"] = np.int64((jp >= 0).sum())
        print(name, prim.size, int((prim >= 0).sum()), f"{time.time() - t0:.1f}s", flush=True)
        del o, desc
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "big_goldens.npz"), **out)


if __name__ == "__main__":
    main()
