#!/usr/bin/env python3
"""Bring-up diagnostics on a GPU box: parity of both precisions against the oracle + quick timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Bvh, EXACT_F64, FAST_F32
from oracle import oracle


def cmp_primary(name, desc):
    o = oracle.OracleScene(desc)
    s = Scene(desc)
    t0 = time.time(); op, ot = o.trace_primary(); to = time.time() - t0
    for prec, label in ((EXACT_F64, "exact"), (FAST_F32, "fast")):
        t0 = time.time(); gp, gt = s.TracePrimary(precision=prec); tg = time.time() - t0
        mism = int((gp != op).sum())
        both = (gp == op) & (op >= 0)
        rel = np.abs(gt[both] - ot[both]) / np.abs(ot[both])
        print(f"[primary] {name:10s} {label:5s} rays={len(op)} hits={(op>=0).sum()} id_mismatch={mism} "
              f"t_exact={bool(np.array_equal(gt, ot))} max_rel_t={rel.max() if len(rel) else 0:.3e} oracle_s={to:.2f} gpu_s={tg:.2f}")
    return o, s


def cmp_hit(name, desc, o, s, n=200000, seed=3):
    rng = np.random.default_rng(seed)
    lo = desc.prims["v"].reshape(-1, 3)
    ext_lo, ext_hi = np.array([-3., -3., -3.]), np.array([3., 3., 3.])
    org = rng.uniform(ext_lo, ext_hi, (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    for tmin, tmax, any_hit in ((1e-6, 99999999., False), (1e-6, 2.0, True)):
        op, osub, ot = o.hit(org, d, tmin, tmax)
        for prec, label in ((EXACT_F64, "exact"), (FAST_F32, "fast")):
            gp, gsub, gt = s.Hit(org, d, tmin, tmax, precision=prec, any_hit=any_hit)
            if any_hit:
                mism = int(((gp >= 0) != (op >= 0)).sum())
                print(f"[hit-any] {name:10s} {label:5s} occluded_mismatch={mism}/{n} occl={(op>=0).sum()}")
            else:
                mism = int((gp != op).sum()); sm = int(((gsub != osub) & (gp == op)).sum())
                print(f"[hit]     {name:10s} {label:5s} id_mismatch={mism}/{n} sub_mismatch={sm} t_exact={bool(np.array_equal(gt, ot))} hits={(op>=0).sum()}")


def cmp_image(name, desc, spp, seed=7):
    o = oracle.OracleScene(desc)
    s = Scene(desc)
    t0 = time.time(); ref, st = o.sample(spp, seed=seed, stats=True); to = time.time() - t0
    out = {}
    for prec, label in ((EXACT_F64, "exact"), (FAST_F32, "fast")):
        integ = CudaPixelIntegrator(s, precision=prec, seed=seed)
        t0 = time.time(); img = integ.Sample(spp).copy(); tg = time.time() - t0
        diff = np.abs(img[:, :, :3] - ref[:, :, :3])
        nbad = int((img[:, :, :3] != ref[:, :, :3]).any(axis=2).sum())
        rmse = float(np.sqrt((diff ** 2).mean())); mean = float(np.abs(ref[:, :, :3]).mean())
        stt = integ.stats
        print(f"[image]   {name:10s} {label:5s} {desc.width}x{desc.height}x{spp} pixels_differ={nbad} max_abs={diff.max():.3e} rmse={rmse:.3e} mean={mean:.3e} "
              f"rays(gpu c/s)={stt['closest_rays']}/{stt['shadow_rays']} rays(orc c/s)={st['closest_rays']}/{st['shadow_rays']} ms={stt['ms_total']:.2f} oracle_s={to:.2f} wall_s={tg:.2f}")
        out[label] = img
    return out


def timing(name, desc, spp, prec, reps=2):
    s = Scene(desc)
    integ = CudaPixelIntegrator(s, precision=prec, seed=1)
    for r in range(reps):
        t0 = time.time(); integ.SampleF32(spp); wall = time.time() - t0
        st = integ.stats
        rays = st["closest_rays"] + st["shadow_rays"]
        print(f"[timing]  {name:10s} {'exact' if prec == EXACT_F64 else 'fast'} {desc.width}x{desc.height}x{spp} rays={rays/1e6:.1f}M ms={st['ms_total']:.1f} "
              f"Mrays/s={rays/st['ms_total']/1e3:.1f} extend_ms={st['ms_extend']:.1f} shadow_ms={st['ms_shadow']:.1f} other_ms={st['ms_shade']:.1f} launches={st['launches']} wall={wall:.2f}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["primary", "hit", "image", "timing"]
    if "primary" in which or "hit" in which:
        for name, kw in (("cornell", {}), ("c1_cube", {}), ("c2_spot", dict(width=480, height=270)), ("c3_renault", dict(width=480, height=270))):
            desc = scenes.WORKLOADS[name](**kw)
            o, s = cmp_primary(name, desc)
            if "hit" in which:
                cmp_hit(name, desc, o, s)
    if "image" in which:
        cmp_image("cornell", scenes.cornell(), 2)
        cmp_image("c1_cube", scenes.c1_cube(width=160, height=120), 4)
        cmp_image("c2_spot", scenes.c2_spot(width=240, height=135), 2)
        cmp_image("c3_renault", scenes.c3_renault(width=160, height=90), 2)
    if "timing" in which:
        timing("cornell", scenes.cornell(), 16, FAST_F32)
        timing("c2_spot", scenes.c2_spot(), 8, FAST_F32, reps=3)
        timing("c2_spot", scenes.c2_spot(), 2, EXACT_F64)
        timing("c3_renault", scenes.c3_renault(), 8, FAST_F32)
