"""Known answers for the sphere sample's oracle (oracle/mafrix_oracle_sky.c restates GetColor of
/root/reference/RenderTest/Sample/RayTracing.fs:367-382).  The reference ships no vectors for it (PARITY UNPINNED):
analytic cases per function, and a second restatement in pure Python that must agree to the last bit."""
import math

import numpy as np
import pytest

from mafrixraytracing_b200 import scenes
from mafrixraytracing_b200.scene import (RayTraceCamera, SceneDesc, SkyTracer, make_materials, sphere_prims, SKY_TRACER)
from oracle import oracle
from . import pyref_sky


def sky_desc(centers, radii, specs, width=8, height=8, max_depth=50, cam=None, tables=(None, None)):
    cam = cam or RayTraceCamera((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, width / height)
    prims = sphere_prims(np.array(centers, float), np.array(radii, float), np.arange(len(specs), dtype=np.int32))
    return SceneDesc(prims, make_materials(specs), None, None, width, height, max_depth, SKY_TRACER,
                     sky=SkyTracer(cam, tables[0], tables[1]))


def as_pyref(desc):
    """The same scene in the Python restatement's own terms."""
    spheres = []
    for p in desc.prims:
        m = desc.materials[p["material"]]
        kind = {0: "lambert", 1: "metal", 3: "dielectric", 4: "checker", 5: "noise"}[int(m["kind"])]
        mat = dict(kind=kind, albedo=tuple(m["albedo"]), fuzz=float(m["fuzz"]), ri=float(m["ei"]), even=tuple(m["albedo"]),
                   odd=(float(m["fuzz"]), float(m["ei"]), float(m["et"])))
        spheres.append((tuple(p["v"][:3]), float(p["v"][3]), mat))
    c = desc.sky.camera
    cam = pyref_sky.lens_camera(tuple(c.lookfrom), tuple(c.lookat), tuple(c.vup), c.vfov, c.aspect, c.aperture, c.focus_dist)
    return spheres, (desc.sky.ranfloat, desc.sky.perm), cam


def test_lens_camera_matches_the_constructor_by_hand():
    # lookfrom (0,0,5) -> lookat origin, vfov 90, aspect 2, focus 5: w = +z, u = +x, v = +y, half_height = tan(45 deg)
    cam = oracle.camera_lens((0, 0, 5), (0, 0, 0), (0, 1, 0), 90.0, 2.0, 0.5, 5.0)
    hh = math.tan(90.0 * math.pi / 180. / 2.)
    assert np.array_equal(cam[0:3], [0, 0, 5])
    assert np.allclose(cam[3:6], [-5 * 2 * hh, -5 * hh, 0.0], rtol=0, atol=1e-15)
    assert np.allclose(cam[6:9], [2 * 5 * 2 * hh, 0, 0]) and np.allclose(cam[9:12], [0, 2 * 5 * hh, 0])
    assert np.array_equal(cam[12:15], [1, 0, 0]) and np.array_equal(cam[15:18], [0, 1, 0]) and cam[18] == 0.25
    # and the Python restatement derives the same 19 numbers to the last bit
    c = pyref_sky.lens_camera((13., 2., 3.), (0., 0., 0.), (0., 1., 0.), 20.0, 2.0, 0.1, 10.0)
    flat = list(c["origin"]) + list(c["lower_left"]) + list(c["horizontal"]) + list(c["vertical"]) + list(c["u"]) + list(c["v"]) + [c["lens_radius"]]
    assert np.array_equal(oracle.camera_lens((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 2.0, 0.1, 10.0), flat)


def test_library_camera_equals_the_oracle_camera():
    cam = RayTraceCamera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 2.0, 0.3, 7.5)
    assert np.array_equal(cam.derived(), oracle.camera_lens((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 2.0, 0.3, 7.5))


def test_sphere_roots_strict_bounds_and_unnormalised_direction():
    o = oracle.OracleSkyScene(sky_desc([(0, 0, 0)], [1.0], [("lambert", (.5, .5, .5))]))
    prim, t = o.hit([(0, 0, 5)], [(0, 0, -3)])              # Ray() normalises: t is a distance
    assert prim[0] == 0 and t[0] == 4.0
    prim, t = o.hit([(0, 0, 0)], [(0, 0, 1)])                # inside: near root negative, far root taken
    assert prim[0] == 0 and t[0] == 1.0
    prim, t = o.hit([(0, 0, 5)], [(0, 0, -1)], tmin=0.00001, tmax=4.0)      # `tmp < tMax` is strict: 4.0 rejected, 6.0 too
    assert prim[0] == -1
    prim, t = o.hit([(0, 0, 5)], [(0, 0, -1)], tmin=4.0, tmax=100.)         # `tmp > tMin` is strict: falls to the far root
    assert prim[0] == 0 and t[0] == 6.0
    prim, t = o.hit([(0, 1, 5)], [(0, 0, -1)])               # tangent: discriminant 0 is a miss (`> 0`)
    assert prim[0] == -1


def test_list_hit_takes_the_nearest_and_the_first_on_ties():
    d = sky_desc([(0, 0, -3), (0, 0, 0), (0, 0, 0)], [1.0, 1.0, 1.0], [("lambert", (.1, .1, .1))] * 3)
    o = oracle.OracleSkyScene(d)
    prim, t = o.hit([(0, 0, 5)], [(0, 0, -1)])
    assert prim[0] == 1 and t[0] == 4.0                      # nearest; spheres 1 and 2 coincide: minBy keeps the first


def test_miss_returns_the_sky_gradient():
    d = sky_desc([(100, 100, 100)], [0.5], [("lambert", (.5, .5, .5))], width=4, height=4)
    o = oracle.OracleSkyScene(d)
    tex = o.sample(1, seed=3)
    spheres, tables, cam = as_pyref(d)
    for (i, j) in [(0, 0), (3, 1), (2, 3)]:
        c = pyref_sky.trace_path(spheres, tables, cam, 4, 4, 50, i, j, 0, 3)
        assert tuple(tex[i, j, :3]) == c
        assert 0.5 <= c[0] <= 1.0 and c[2] == 1.0            # (1-t) + t*(0.5, 0.7, 1.0)


def test_depth_limit_returns_black_and_counts_like_the_reference():
    # a sphere that fills the view: with max_depth 0 the hit at depth 0 fails `depth < max_depth` -> Color()
    kw = dict(width=2, height=2, cam=RayTraceCamera((0, 0, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.0))
    o = oracle.OracleSkyScene(sky_desc([(0, 0, 0)], [2.0], [("lambert", (.9, .9, .9))], max_depth=0, **kw))
    tex, st = o.sample(2, seed=1, stats=True)
    assert np.all(tex[:, :, :3] == 0.0) and st["closest_rays"] == 2 * 2 * 2
    # max_depth 1: the scattered ray leaves the (convex, outward-facing) sphere and returns albedo * sky
    o = oracle.OracleSkyScene(sky_desc([(0, 0, 0)], [2.0], [("lambert", (.9, .9, .9))], max_depth=1, **kw))
    tex, st = o.sample(2, seed=1, stats=True)
    assert st["closest_rays"] == 2 * 2 * 2 * 2
    assert np.all(tex[:, :, :3] > 0.9 * 0.5 - 1e-12) and np.all(tex[:, :, :3] <= 0.9)


MIXED = dict(centers=[(0, -100.5, -1), (0, 0, -1), (-1.05, 0, -1), (1.05, 0, -1), (0.3, 0.9, -0.6)],
             radii=[100.0, 0.5, 0.5, 0.5, 0.3],
             specs=[("checker", (0.2, 0.3, 0.1), (0.9, 0.9, 0.9)), ("lambert", (0.8, 0.3, 0.3)), ("dielectric", 1.5),
                    ("metal", (0.8, 0.6, 0.2), 0.3), ("noise",)])


@pytest.mark.parametrize("aperture", [0.0, 0.4])
def test_get_color_matches_the_python_restatement_bit_for_bit(aperture):
    rf, pm = scenes.perlin_tables(3)
    cam = RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, 1.5, aperture, 3.4)
    d = sky_desc(MIXED["centers"], MIXED["radii"], MIXED["specs"], width=12, height=8, max_depth=50, cam=cam, tables=(rf, pm))
    o = oracle.OracleSkyScene(d)
    spheres, tables, pcam = as_pyref(d)
    kinds = set()
    for (i, j, s) in [(x, y, s) for x in range(0, 12, 2) for y in range(0, 8, 2) for s in range(3)]:
        ref = pyref_sky.trace_path(spheres, tables, pcam, 12, 8, 50, i, j, s, 11)
        got = o.trace_path(i, j, s, seed=11)
        assert tuple(got) == ref, (i, j, s)
        kinds.add(ref != (0., 0., 0.))
    assert kinds == {True} or kinds == {True, False}
    # the frame is the per-pixel mean of those paths in sample order
    tex = o.sample(3, seed=11)
    acc = (0., 0., 0.)
    for s in range(3):
        c = pyref_sky.trace_path(spheres, tables, pcam, 12, 8, 50, 4, 6, s, 11)
        acc = (acc[0] + c[0], acc[1] + c[1], acc[2] + c[2])
    assert tuple(tex[4, 6, :3]) == (acc[0] / 3., acc[1] / 3., acc[2] / 3.) and tex[4, 6, 3] == 1.0


def test_every_material_is_exercised_by_the_mixed_scene():
    rf, pm = scenes.perlin_tables(3)
    d = sky_desc(MIXED["centers"], MIXED["radii"], MIXED["specs"], width=48, height=32,
                 cam=RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, 1.5), tables=(rf, pm))
    prim, _ = oracle.OracleSkyScene(d).trace_primary()
    assert set(np.unique(prim)) >= {0, 1, 2, 3, 4}


def test_random_scene_recipe():
    d = scenes.random_scene(width=40, height=20)
    n = len(d.prims)
    assert 150 <= n <= 173 and np.all(d.prims["kind"] == 2)
    assert tuple(d.prims["v"][n - 4][:4]) == (0., -1000., 0., 1000.) and d.materials["kind"][n - 4] == 5
    assert list(d.materials["kind"][n - 3:]) == [3, 0, 1]
    small = d.prims["v"][:n - 4]
    assert np.all(small[:, 3] == 0.2) and np.all(small[:, 1] == 0.2)
    assert np.all(np.sqrt(((small[:, :3] - [4, 0.2, 0]) ** 2).sum(1)) > 0.9)
    frac = np.bincount(d.materials["kind"][:n - 4], minlength=4) / (n - 4)
    assert 0.65 < frac[0] < 0.92 and frac[2] == 0.0          # no SpecularTransmission in the sphere sample
    rf, pm = d.sky.ranfloat, d.sky.perm.reshape(3, 256)
    assert all(sorted(p) == list(range(256)) for p in pm) and 0.0 <= rf.min() and rf.max() < 1.0
    img = oracle.OracleSkyScene(d).sample(2, seed=1)
    assert img[:, :, :3].min() >= 0.0 and 0.2 < img[:, :, :3].mean() < 0.9


def _sky_goldens():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sky_goldens.npz"))


def golden_sky_scenes():
    rf, pm = scenes.perlin_tables(3)
    cam = RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, 1.5, 0.4, 3.4)
    return {"random_scene": (scenes.random_scene(), scenes.random_scene(width=120, height=60)),
            "mixed": (sky_desc(MIXED["centers"], MIXED["radii"], MIXED["specs"], width=96, height=64, cam=cam, tables=(rf, pm)),) * 2}


@pytest.mark.parametrize("name", ["random_scene", "mixed"])
def test_oracle_matches_the_committed_sky_goldens(name):
    """tests/golden/sky_goldens.npz (minted by tests/golden/make_goldens.py sky) guards the restatement against drift."""
    import hashlib
    g = _sky_goldens()
    full, small = golden_sky_scenes()[name]
    prim, t = oracle.OracleSkyScene(full).trace_primary()
    assert np.array_equal(prim, g[f"primary/{name}/prim"]) and np.array_equal(t[::97], g[f"primary/{name}/t_stride97"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(t.tobytes()).digest(), dtype=np.uint8), g[f"primary/{name}/sha_t"])
    tex = oracle.OracleSkyScene(small).sample(2, seed=7)
    assert np.array_equal(tex[:, :, :3], g[f"image/{name}/rgb"])
    disp = oracle.sky_display_rgba8(tex)
    assert np.array_equal(np.frombuffer(hashlib.sha256(disp.tobytes()).digest(), dtype=np.uint8), g[f"image/{name}/display_sha"])


def test_display_transform_known_answers():
    # sqrt, int(255.99 c) (truncation), row j of the texture on screen row h-1-j (RayTracing.fs:456-460)
    tex = np.zeros((2, 3, 4))
    tex[0, 0, :3] = (1.0, 0.25, 0.0)
    tex[1, 2, :3] = (0.5, 0.04, 1.0 / 255.99 ** 2)
    out = oracle.sky_display_rgba8(tex)
    assert out.shape == (3, 2, 4)
    assert tuple(out[2, 0]) == (255, 127, 0, 255)                    # texture [0,0] -> bottom-left
    assert tuple(out[0, 1][:2]) == (int(255.99 * math.sqrt(0.5)), int(255.99 * 0.2)) and out[0, 1, 2] in (0, 1)
