"""Known-answer tests that pin the CPU oracle (the reference ships none):
analytic hits, the reference's quirks Q1-Q7/Q10, the RNG's published vectors, an independent
pure-Python restatement of the integrator, and the committed golden vectors."""
import hashlib
import math

import numpy as np
import pytest

from mafrixraytracing_b200 import scenes
from mafrixraytracing_b200.scene import (AreaLight, PinholeCamera, SceneDesc, make_materials, make_prims,
                                         rect_prim, sphere_prims, TRIANGLE)
from oracle import oracle
from tests import pyref


def tri(v0, v1, v2, material=0):
    p = make_prims(1)
    p["kind"] = TRIANGLE
    p["material"] = material
    p["v"][0, :9] = np.concatenate([v0, v1, v2])
    return p


def scene_of(prims, width=8, height=8, max_depth=0):
    mats = make_materials([("lambert", (0.5, 0.5, 0.5))])
    light = AreaLight(np.array([(-1, 5, 1), (-1, 5, -1), (1, 5, -1), (1, 5, 1)], float), (0, -1, 0), (10, 10, 10))
    cam = PinholeCamera((0, 0, 5), (0, 0, -1), 120.0, 1.0)
    return oracle.OracleScene(SceneDesc(np.concatenate(prims), mats, light, cam, width, height, max_depth))


UNIT = tri((0, 0, 0), (1, 0, 0), (0, 1, 0))
BIG = 99999999.


def one(o, origin, d, tmin=1e-6, tmax=BIG):
    prim, sub, t = o.hit([origin], [d], tmin, tmax)
    return int(prim[0]), int(sub[0]), float(t[0])


# ---------------------------------------------------------------- RNG
def test_philox_published_vectors():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in oracle.philox([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    assert pyref.philox4x32_10([0] * 4, [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


# ---------------------------------------------------------------- Triangle (Trangle.fs:120-155)
def test_triangle_known_t_and_two_sided():
    o = scene_of([UNIT])
    assert one(o, (0.25, 0.25, 1), (0, 0, -1)) == (0, 0, 1.0)
    assert one(o, (0.25, 0.25, -2), (0, 0, 1)) == (0, 0, 2.0)          # back face is hit too
    assert one(o, (2, 2, 1), (0, 0, -1))[0] == -1


def test_triangle_parallel_and_tmin():
    o = scene_of([UNIT])
    assert one(o, (-1, 0.25, 0), (1, 0, 0))[0] == -1                    # |divisor| < 1e-6
    assert one(o, (0.25, 0.25, 0), (0, 0, -1))[0] == -1                 # t = 0 <= tMin
    assert one(o, (0.25, 0.25, 5e-7), (0, 0, -1))[0] == -1              # t = 5e-7 <= 1e-6
    assert one(o, (0.25, 0.25, 2e-6), (0, 0, -1))[0] == 0


def test_triangle_edge_rules():
    o = scene_of([UNIT])
    # exact edge hits need an (unnormalised, exactly representable) slanted direction: with a
    # zero direction component the box test sees 0/0 = NaN and rejects (next test)
    assert one(o, (0.5, 0.3, 1), (-0.5, 0, -1))[:2] == (0, 0)           # b1 == 0 accepted
    assert one(o, (0.3, 0.5, 1), (0, -0.5, -1))[:2] == (0, 0)           # b2 == 0 accepted
    assert one(o, (0.5, 0.5, 1), (0, 0, -1))[0] == -1                   # b1 + b2 == 1 rejected (>= 1.)
    assert one(o, (1.0, 0.0, 1), (0, 0, -1))[0] == -1                   # vertex v1: b1 = 1, b1+b2 = 1


def test_origin_on_box_plane_with_zero_direction_is_nan_miss_quirk_q6():
    # (pmin.x - o.x)/dir.x = 0/0 = NaN survives the selects and fails `tmin < tMax` (IHitable.fs:54)
    o = scene_of([UNIT])
    assert one(o, (0.0, 0.3, 1), (0, 0, -1))[0] == -1
    assert one(o, (1e-9, 0.3, 1), (0, 0, -1))[0] == 0


def test_triangle_ignores_tmax_quirk_q2():
    # tilted triangle: box entry at t=1 (< tMax) but the hit is at t=1.9 (> tMax); a single-primitive
    # leaf returns it anyway because Triangle.Hit never looks at tMax (Trangle.fs:148, BvhNode.fs:76-80)
    t1 = tri((0, 0, 0), (1, 0, 1), (0, 1, 0))
    o = scene_of([t1])
    prim, _, t = one(o, (0.1, 0.1, 2), (0, 0, -1), 1e-6, 1.5)
    assert prim == 0 and abs(t - 1.9) < 1e-12
    # a missing sibling in the same leaf has key tMax = 1.5 < 1.9 and wins -> no hit
    o2 = scene_of([t1, tri((5, 5, 0), (6, 5, 0), (5, 6, 0))])
    assert one(o2, (0.1, 0.1, 2), (0, 0, -1), 1e-6, 1.5)[0] == -1


# ---------------------------------------------------------------- Rect (Rect.fs:26-31)
def test_rect_returns_tri1_even_if_tri2_is_nearer_quirk_q3():
    folded = rect_prim((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0.2, 0.5))
    o = scene_of([folded])
    prim, sub, t = one(o, (0.8, 0.4, 2), (0, 0, -1))
    assert (prim, sub, t) == (0, 0, 2.0)                                # tri2 would be hit at t < 2


def test_rect_second_triangle():
    q = rect_prim((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0))
    o = scene_of([q])
    assert one(o, (0.8, 0.2, 1), (0, 0, -1))[:2] == (0, 0)
    assert one(o, (0.2, 0.8, 1), (0, 0, -1))[:2] == (0, 1)


# ---------------------------------------------------------------- Sphere (Sphere.fs:21-43)
def test_sphere_roots_and_tmax():
    o = scene_of([sphere_prims([(0, 0, 0)], 1.0, 0)])
    assert one(o, (0, 0, 3), (0, 0, -1)) == (0, 0, 2.0)
    assert one(o, (0, 0, 0), (0, 0, -1)) == (0, 0, 1.0)                 # inside: far root
    assert one(o, (0, 0, 3), (0, 0, -1), 1e-6, 1.5)[0] == -1            # honours tMax
    assert one(o, (1, 0, 3), (0, 0, -1))[0] == -1                       # tangent: discriminant == 0 is a miss
    prim, sub, t, pt, nm = o.hit([(0, 0, 3)], [(0, 0, -1)], 1e-6, BIG, want_geometry=True)
    assert np.array_equal(pt[0], [0, 0, 1]) and np.array_equal(nm[0], [0, 0, 1])


# ---------------------------------------------------------------- AABB (IHitable.fs:18-54)
def test_aabb_zero_direction_and_negative_zero_quirk_q6():
    o = scene_of([UNIT])
    assert one(o, (0.25, 0.25, 2), (0.0, 0.0, -1.0))[0] == 0           # +0: slabs are +-inf, accepted
    assert one(o, (1.5, 0.25, 2), (0.0, 0.0, -1.0))[0] == -1           # outside the x slab
    # `dir.x >= 0.` is true for -0.0, so the slab is taken in the wrong order: (+inf, -inf) -> rejected
    assert one(o, (0.25, 0.25, 2), (-0.0, 0.0, -1.0))[0] == -1


# ---------------------------------------------------------------- Bvh (BvhNode.fs:24-83)
def test_bvh_structure_median_split_heap_index():
    v, f = scenes.load_mesh("spot")
    d = scenes.c2_spot(width=8, height=8)
    o = oracle.OracleScene(d)
    nodes, idx = o.bvh()
    n = len(d.prims)
    assert len(nodes) == 2 * n - 1 and sorted(idx.tolist()) == list(range(n))
    assert nodes[0]["first"] == 0 and nodes[0]["count"] == n
    todo, leaves = [0], 0
    while todo:
        i = todo.pop()
        nd = nodes[i]
        if nd["count"] > 3:
            l, r = nodes[2 * i + 1], nodes[2 * i + 2]
            assert l["count"] == nd["count"] // 2 and r["count"] == nd["count"] - nd["count"] // 2
            assert l["first"] == nd["first"] and r["first"] == nd["first"] + l["count"]
            assert (np.minimum(l["pmin"], r["pmin"]) == nd["pmin"]).all() and (np.maximum(l["pmax"], r["pmax"]) == nd["pmax"]).all()
            todo += [2 * i + 1, 2 * i + 2]
        else:
            leaves += 1
    assert leaves >= n / 3


def test_tie_rules_quirk_q1():
    # 8 coincident triangles: leaves [0,1][2,3][4,5][6,7]; each leaf keeps its FIRST minimum
    # (Array.minBy), each interior node prefers the RIGHT child on equal t (BvhNode.fs:69-70) -> 6
    o = scene_of([UNIT] * 8)
    nodes, idx = o.bvh()
    assert idx.tolist() == list(range(8))                               # stable sort keeps equal keys in order
    assert one(o, (0.25, 0.25, 1), (0, 0, -1)) == (6, 0, 1.0)
    o3 = scene_of([UNIT] * 3)                                           # single leaf: first wins
    assert one(o3, (0.25, 0.25, 1), (0, 0, -1)) == (0, 0, 1.0)


def test_exhaustive_equals_brute_force():
    d = scenes.c1_cube(width=8, height=8)
    o = oracle.OracleScene(d)
    rng = np.random.default_rng(5)
    org = rng.uniform(-2.9, 2.9, (300, 3))
    dr = rng.normal(size=(300, 3))
    dr /= np.linalg.norm(dr, axis=1)[:, None]
    prim, sub, t = o.hit(org, dr, 1e-6, BIG)
    for r in range(300):
        best = (np.inf, -1)
        for i, p in enumerate(d.prims):
            v = p["v"].reshape(4, 3)
            tris = [(v[0], v[1], v[2])] + ([(v[0], v[2], v[3])] if p["kind"] == 1 else [])
            for (a, b, c) in tris:
                tt = pyref.tri_hit(tuple(a), tuple(b), tuple(c), tuple(org[r]), tuple(dr[r]), 1e-6)
                if tt is not None:
                    if tt < best[0]:
                        best = (tt, i)
                    break
        assert prim[r] == best[1] and (best[1] < 0 or t[r] == best[0])


# ---------------------------------------------------------------- Camera (Camera.fs:96-139)
def test_camera_effective_fov_is_half_quirk_q5():
    cam = oracle.camera_pinhole((0, 0, 0), (0, 0, -1), 120.0, 1.0)
    h = math.tan(0.5 * 120.0 * math.pi / 360.)
    assert np.allclose(cam[6:9], [h, 0, 0]) and np.allclose(cam[9:12], [0, -h, 0])
    assert np.allclose(cam[3:6], [-h / 2, h / 2, -0.5])
    # half-angle of the frustum = atan((h/2)/0.5) = 30 deg -> full FOV 60 = fov/2
    assert abs(math.degrees(math.atan(h)) - 30.0) < 1e-12
    cam2 = oracle.camera_pinhole((0, 0, 0), (0, 0, -3), 120.0, 2.0)     # dir is normalised, v = h/aspect
    assert np.allclose(cam2[9:12], [0, -h / 2, 0])


# ---------------------------------------------------------------- integrator vs the pure-Python restatement
def _tiny():
    d = scenes.cornell(width=12, height=12, max_depth=2)
    rects = [tuple(tuple(p["v"][3 * k:3 * k + 3]) for k in range(4)) + (int(p["material"]),) for p in d.prims]
    alb = [tuple(m["albedo"]) for m in d.materials]
    cam = oracle.camera_pinhole(d.camera.pos_arg, d.camera.dir_arg, d.camera.fov, d.camera.aspect)
    ts = pyref.TinyScene(rects, alb, [tuple(p) for p in d.light.p], tuple(d.light.normal), tuple(d.light.color),
                         list(cam), d.width, d.height, d.max_depth)
    return d, ts


def test_path_integrator_matches_python_restatement():
    d, ts = _tiny()
    o = oracle.OracleScene(d)
    nonzero = 0
    for (i, j, s) in [(2, 3, 0), (6, 6, 1), (9, 4, 2), (5, 10, 0), (7, 8, 5), (3, 9, 1), (10, 10, 3), (6, 2, 4)]:
        got = o.trace_path(i, j, s, seed=11)
        want = ts.trace_path(i, j, s, 11)
        assert np.array_equal(got, np.array(want)), (i, j, s, got, want)
        nonzero += int(any(want))
    assert nonzero >= 4


def test_depth_counts_down_to_minus_one_quirk_q10():
    d = scenes.cornell(width=6, height=6, max_depth=0)
    o = oracle.OracleScene(d)
    tex, st = o.sample(1, seed=3, stats=True)
    # D=0: one shaded vertex per hit path, and one wasted closest query at depth -1 per shaded vertex
    assert st["shadow_rays"] == st["wasted_rays"] > 0 and st["closest_rays"] == 36


def test_sample_layout_is_x_major_color_wh():
    d = scenes.cornell(width=10, height=6, max_depth=1)
    o = oracle.OracleScene(d)
    tex = o.sample(2, seed=5)
    assert tex.shape == (10, 6, 4) and (tex[:, :, 3] == 1.0).all()
    x, y = 7, 2
    want = (o.trace_path(x, y, 0, 5) + o.trace_path(x, y, 1, 5)) / 2.0
    assert np.array_equal(tex[x, y, :3], want)
    part = np.zeros_like(tex)
    o.sample(2, seed=5, region=(4, 1, 9, 5), out=part)
    assert np.array_equal(part[4:9, 1:5], tex[4:9, 1:5]) and not part[:4].any()


# ---------------------------------------------------------------- Film + ACES (Film.fs:18-23, Scene.fs:273-330)
def test_film_and_tonemap():
    rng = np.random.default_rng(1)
    f1, f2 = rng.random((5, 4, 4)), rng.random((5, 4, 4))
    f1[..., 3] = f2[..., 3] = 1.0
    s = np.zeros_like(f1)
    t1 = oracle.film_add_sample(s, f1, 1.0)
    assert np.array_equal(t1[..., :3], f1[..., :3])
    t2 = oracle.film_add_sample(s, f2, 2.0)
    assert np.array_equal(t2[..., :3], (f1[..., :3] + f2[..., :3]) / 2.0)
    tex = np.zeros((3, 2, 4))
    tex[0, 0, :3] = (0.0, 0.18, 1.0)
    tex[2, 1, :3] = (5.0, -1.0, 0.5)
    out = oracle.tonemap_rgba8(tex)

    def aces(x):
        q = (x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14)
        q = min(max(q, 0.0), 1.0)
        return int(255.99 * math.sqrt(q))
    assert out.shape == (2, 3, 4)
    assert out[0, 0].tolist() == [aces(0.0), aces(0.18), aces(1.0), 255]
    assert out[1, 2].tolist() == [aces(5.0), aces(-1.0), aces(0.5), 255]


# ---------------------------------------------------------------- committed goldens
def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("name,kw", [("cornell", {}), ("c1_cube", {}), ("c2_spot_small", dict(width=480, height=270)),
                                     ("c3_renault_small", dict(width=480, height=270))])
def test_oracle_primary_matches_goldens(goldens, name, kw):
    o = oracle.OracleScene(scenes.WORKLOADS[name.replace("_small", "")](**kw))
    prim, t = o.trace_primary()
    assert np.array_equal(prim, goldens[f"primary/{name}/prim"])
    assert np.array_equal(_sha(t), goldens[f"primary/{name}/sha_t"])
    assert np.array_equal(t[::97], goldens[f"primary/{name}/t_stride97"])


@pytest.mark.parametrize("name", ["cornell", "c1_cube", "c2_spot", "c3_renault", "spheres"])
def test_oracle_images_match_goldens(goldens, name):
    from tests.golden.make_goldens import IMAGES, builder
    kw, spp, seed = IMAGES[name]
    o = oracle.OracleScene(builder(name)(**kw))
    assert np.array_equal(o.sample(spp, seed=seed)[:, :, :3], goldens[f"image/{name}/rgb"])


def test_ordered_counts_are_below_exhaustive():
    o = oracle.OracleScene(scenes.c2_spot(width=48, height=27))
    _, st = o.sample(1, seed=2, stats=True, count_ordered=True)
    rays = st["closest_rays"] + st["wasted_rays"] + st["shadow_rays"]
    assert st["ord_rays"] == [st["closest_rays"], st["shadow_rays"]]
    assert sum(st["ord_nodes"]) < st["ref_nodes"] and st["ref_nodes"] / rays > 20
