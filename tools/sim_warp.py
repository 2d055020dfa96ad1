#!/usr/bin/env python3
"""CPU model of the persistent-warp traversal loop (k_f_trace4 in csrc/mfx_fast.cu): 32 lanes in
lockstep, same state machine (refill -> node step -> leaf vote -> pop), on the same flattened
layout (children pairs indexed by heap index, leaf-order slots).  Used to debug the control
logic without a GPU and to check the stackless bit-trail traversal against the oracle
(tests/test_traversal_model.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Bvh

FULL = 0xffffffff


def flatten(desc):
    b = Bvh.Build(desc.prims)
    nodes, idx = b.nodes, b.indices
    prims = desc.prims
    slots, ffirst = [], []
    for s in range(len(idx)):
        p = prims[idx[s]]
        ffirst.append(len(slots))
        v = p["v"].reshape(4, 3)
        if p["kind"] == 2:
            slots.append(("s", v[0].copy(), float(p["v"][3]), s))
        else:
            for k in range(2 if p["kind"] == 1 else 1):
                slots.append(("t", v[0].copy(), v[1 + k] - v[0], v[2 + k] - v[0], s))
    ffirst.append(len(slots))

    def meta(nd):
        f0, f1 = ffirst[nd["first"]], ffirst[nd["first"] + nd["count"]]
        return (f0 << 3) | (f1 - f0)
    pairs = {}
    root = nodes[0]
    root_meta = meta(root) if root["count"] <= 3 else -1
    todo = [0] if root_meta < 0 else []
    while todo:
        i = todo.pop()
        L, R = nodes[2 * i + 1], nodes[2 * i + 2]
        li, ri = L["count"] > 3, R["count"] > 3
        pairs[i + 1] = (L["pmin"].astype(np.float32), L["pmax"].astype(np.float32), R["pmin"].astype(np.float32),
                        R["pmax"].astype(np.float32), -1 if li else meta(L), -1 if ri else meta(R))
        if li: todo.append(2 * i + 1)
        if ri: todo.append(2 * i + 2)
    return dict(pairs=pairs, slots=slots, root=(root["pmin"].astype(np.float32), root["pmax"].astype(np.float32)),
                root_meta=root_meta, ref=idx)


def box(o, idir, lo, hi, tmin, tmax):
    t0 = (lo - o) * idir
    t1 = (hi - o) * idir
    tn = max(np.minimum(t0, t1).max(), tmin)
    tf = min(np.maximum(t0, t1).min(), tmax)
    return tn <= tf, tn


def leaf(sc, o, d, meta, best_t, best_slot, tmin=1e-6):
    first, cnt = meta >> 3, meta & 7
    found = False
    for k in range(cnt):
        sl = sc["slots"][first + k]
        if sl[0] == "t":
            _, v0, e1, e2, prim = sl
            s1 = np.cross(d, e2)
            div = s1 @ e1
            if abs(div) < 1e-6: continue
            inv = 1.0 / div
            dd = o - v0
            b1 = (dd @ s1) * inv
            if b1 < 0 or b1 > 1: continue
            s2 = np.cross(dd, e1)
            b2 = (d @ s2) * inv
            if b2 < 0 or b1 + b2 >= 1: continue
            t = (e2 @ s2) * inv
            if t > tmin and t < best_t:
                best_t, best_slot, found = t, first + k, True
    return found, best_t, best_slot


def clz(x):
    return 32 - int(x).bit_length()


def run_warp(sc, rays, REFILL_T=12, LEAF_T=16, ANY=False, max_iters=200000):
    """rays: list of (origin f32[3], dir f32[3], tmax).  Returns ({ray: (t, slot)}, iterations, stuck-state or None)."""
    n = len(rays)
    cursor = 0
    L = 32
    pid = [-1] * L; o = [None] * L; d = [None] * L; idir = [None] * L
    best_t = [0.0] * L; best_slot = [-1] * L
    h = [1] * L; pend = [0] * L; depth = [0] * L; needPop = [False] * L
    leafA = [-1] * L; leafB = [-1] * L; eB = [0.0] * L
    entry = [dict() for _ in range(L)]
    exhausted = False
    out = {}
    iters = 0
    while True:
        iters += 1
        if iters > max_iters:
            return out, iters, dict(pid=pid, leafA=leafA, needPop=needPop, pend=pend, h=h, exhausted=exhausted, cursor=cursor)
        idle = sum(1 << l for l in range(L) if pid[l] < 0)
        idle_now = idle
        if not exhausted and bin(idle).count("1") >= REFILL_T:
            nidle = bin(idle).count("1")
            base = cursor; cursor += nidle
            if base + nidle >= n: exhausted = True
            for l in range(L):
                if pid[l] < 0:
                    idx = base + bin(idle & ((1 << l) - 1)).count("1")
                    if idx < n:
                        pid[l] = idx
                        o[l], d[l], tmax = rays[idx]
                        with np.errstate(divide="ignore"):
                            idir[l] = (1.0 / d[l]).astype(np.float32)
                        best_t[l] = tmax; best_slot[l] = -1; pend[l] = 0; leafA[l] = leafB[l] = -1; h[l] = 1; depth[l] = 0
                        inb, _ = box(o[l], idir[l], sc["root"][0], sc["root"][1], 1e-6, best_t[l])
                        needPop[l] = (not inb) or sc["root_meta"] >= 0
                        if inb and sc["root_meta"] >= 0: leafA[l] = sc["root_meta"]
            idle_now = sum(1 << l for l in range(L) if pid[l] < 0)
        if idle_now == FULL:
            if exhausted: break
            continue
        for l in range(L):                                   # node step
            if pid[l] >= 0 and leafA[l] < 0 and not needPop[l]:
                Lmin, Lmax, Rmin, Rmax, mL, mR = sc["pairs"][h[l]]
                hitL, eL = box(o[l], idir[l], Lmin, Lmax, 1e-6, best_t[l])
                hitR, eR = box(o[l], idir[l], Rmin, Rmax, 1e-6, best_t[l])
                rFirst = eR < eL
                lfL, lfR = hitL and mL >= 0, hitR and mR >= 0
                inL, inR = hitL and mL < 0, hitR and mR < 0
                leafA[l] = ((mR if (lfR and rFirst) else mL) if lfL else (mR if lfR else -1))
                leafB[l] = (mL if rFirst else mR) if (lfL and lfR) else -1
                eB[l] = eL if rFirst else eR
                both = inL and inR
                rNear = inR and ((not inL) or rFirst)
                if inL or inR: depth[l] += 1
                if both:
                    pend[l] |= 1 << depth[l]
                    entry[l][depth[l]] = eL if rNear else eR
                needPop[l] = not (inL or inR)
                if not needPop[l]: h[l] = 2 * h[l] + (1 if rNear else 0)
        fin = [False] * L
        lp = sum(1 << l for l in range(L) if pid[l] >= 0 and leafA[l] >= 0)
        if lp and (bin(lp).count("1") >= LEAF_T or ((~idle_now & ~lp) & FULL) == 0):     # leaf phase by vote
            for l in range(L):
                if pid[l] >= 0 and leafA[l] >= 0:
                    found, best_t[l], best_slot[l] = leaf(sc, o[l], d[l], leafA[l], best_t[l], best_slot[l])
                    if leafB[l] >= 0 and not (ANY and found) and eB[l] <= best_t[l]:
                        f2, best_t[l], best_slot[l] = leaf(sc, o[l], d[l], leafB[l], best_t[l], best_slot[l])
                        found = found or f2
                    leafA[l] = leafB[l] = -1
                    if ANY and found: fin[l] = True
        for l in range(L):                                   # pop
            if pid[l] >= 0 and needPop[l] and leafA[l] < 0 and not fin[l]:
                while True:
                    if pend[l] == 0: fin[l] = True; break
                    b = 31 - clz(pend[l])
                    pend[l] ^= 1 << b
                    if ANY or entry[l][b] <= best_t[l]:
                        assert depth[l] >= b
                        h[l] = (h[l] >> (depth[l] - b)) ^ 1
                        depth[l] = b
                        needPop[l] = False
                        break
        for l in range(L):
            if fin[l]:
                out[pid[l]] = (best_t[l], best_slot[l])
                pid[l] = -1
    return out, iters, None


def camera_rays(desc, seed=None, uv_out=None):
    """Pixel-centre rays, or jittered ones (seed given): centres sit exactly on quad diagonals."""
    cam = desc.camera
    rng = np.random.default_rng(seed) if seed is not None else None
    rays = []
    for j in range(desc.height):
        for i in range(desc.width):
            ju, jv = (rng.random(), rng.random()) if rng is not None else (0.5, 0.5)
            if uv_out is not None: uv_out.append(((i + ju) / desc.width, (j + jv) / desc.height))
            oo, dd = cam.GetRay((i + ju) / desc.width, (j + jv) / desc.height)
            rays.append((oo.astype(np.float32), dd.astype(np.float32), 99999999.0))
    return rays


if __name__ == "__main__":
    desc = scenes.c2_spot(width=24, height=14)
    sc = flatten(desc)
    rays = camera_rays(desc)
    for rt, lt in ((1, 1), (12, 16), (1, 32)):
        out, iters, stuck = run_warp(sc, rays, rt, lt)
        print(f"REFILL_T={rt} LEAF_T={lt}: iters={iters} finished={len(out)}/{len(rays)} stuck={stuck is not None}")
