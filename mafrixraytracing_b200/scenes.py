"""The BASELINE workloads as SceneDesc builders -- synthetic inputs for the
tests and bench.py.  Meshes come from the committed binary fixtures tests/golden/meshes/*.npz
(minted from the reference's 3DModel/*.obj by tests/golden/make_meshes.py); /root/reference is
never read at run time.

  cornell()   the reference's own default scene (Scene.xml / RayTracing4.fs:9-72) re-authored:
              CornellBox-Original geometry as Rects, 300x300, fov 120, maxDepth 3 (Scene.fs:304)
  c1_cube()   C1: Cube.obj inside a 5-quad open box, 640x480, 16 spp, D=5, PathIntegrator
  c2_spot()   C2: spot on a floor + back wall, 1920x1080, 64 spp, D=5, PathIntegrator
  c3_renault()C3: Renault12TL, mixed materials, NewPathTracer, 1920x1080, 256 spp
  c4_spheres()C4: ~100k spheres (RayTracing.fs:384-415 recipe scaled), 3840x2160, 128 spp
  c5_soup()   C5: spot instanced 1708x (10.0M triangles), 3840x2160, 64 spp
  random_scene()  the sphere sample itself (RenderTest/Sample/RayTracing.fs:384-433): RandomScene + RayTraceCamera,
              400x200, 9 spp, depth 50, rendered by GetColor (SKY_TRACER)

Surfaces are oriented deliberately: the reference never flips normals toward the ray
(quirk Q4), so rooms face inward and objects outward.
"""
import os

import numpy as np

from .scene import (AreaLight, PinholeCamera, RayTraceCamera, SceneDesc, SkyTracer, make_materials, make_prims,
                    rect_prim, sphere_prims, triangles_from_mesh, NEW_PATH_TRACER, PATH_INTEGRATOR, SKY_TRACER, RECT,
                    TRIANGLE)

_HERE = os.path.dirname(os.path.abspath(__file__))
MESH_DIR = os.environ.get("MFX_MESH_DIR", os.path.join(_HERE, "..", "tests", "golden", "meshes"))

WHITE, GREEN, RED = (0.725, 0.71, 0.68), (0.14, 0.45, 0.091), (0.63, 0.065, 0.05)   # Scene.xml:15,18,21


def load_mesh(name):
    z = np.load(os.path.join(MESH_DIR, name + ".npz"))
    return z["v"], z["f"]


def _normal(p0, p1, p2):
    return np.cross(np.asarray(p1, float) - np.asarray(p0, float), np.asarray(p2, float) - np.asarray(p0, float))


def quad_facing(p0, p1, p2, p3, want, material=0):
    """A Rect whose geometric normal (v1-v0)x(v2-v0) points along `want`."""
    if np.dot(_normal(p0, p1, p2), want) < 0:
        p1, p3 = p3, p1
    return rect_prim(p0, p1, p2, p3, material)


def _light(p0, p1, p2, p3, color):
    return AreaLight(np.array([p0, p1, p2, p3], float), np.array([0., -1., 0.]), np.array(color, float))


def _box_from_top(top, material, with_top=True):
    """Side quads (down to y=0) + top of a box given its 4 top corners; outward normals."""
    top = [np.array(t, float) for t in top]
    c = sum(top) / 4.0
    prims = []
    if with_top:
        prims.append(quad_facing(top[0], top[1], top[2], top[3], (0, 1, 0), material))
    for k in range(4):
        a, b = top[k], top[(k + 1) % 4]
        a0, b0 = a.copy(), b.copy()
        a0[1] = 0.0
        b0[1] = 0.0
        mid = (a + b) / 2.0
        out = mid - c
        out[1] = 0.0
        prims.append(quad_facing(a0, a, b, b0, out, material))
    return prims


def cornell(width=300, height=300, max_depth=3, integrator=PATH_INTEGRATOR):
    """Scene.xml re-authored.  CornellBox-Original.obj is absent from the reference repo
    (Scene.xml:10, .gitignore:73); the geometry below is that model's published quads and the
    light quad is the one hard-coded at Scene.fs:194."""
    P = []
    P.append(quad_facing((-1.01, 0, 0.99), (1, 0, 0.99), (1, 0, -1.04), (-0.99, 0, -1.04), (0, 1, 0), 0))       # floor
    P.append(quad_facing((-1.02, 1.99, 0.99), (-1.02, 1.99, -1.04), (1, 1.99, -1.04), (1, 1.99, 0.99), (0, -1, 0), 0))  # ceiling
    P.append(quad_facing((-0.99, 0, -1.04), (1, 0, -1.04), (1, 1.99, -1.04), (-1.02, 1.99, -1.04), (0, 0, 1), 0))       # backWall
    P.append(quad_facing((1, 0, -1.04), (1, 0, 0.99), (1, 1.99, 0.99), (1, 1.99, -1.04), (-1, 0, 0), 1))        # rightWall (green)
    P.append(quad_facing((-1.01, 0, 0.99), (-0.99, 0, -1.04), (-1.02, 1.99, -1.04), (-1.02, 1.99, 0.99), (1, 0, 0), 2))  # leftWall (red)
    P += _box_from_top([(0.53, 0.6, 0.75), (0.70, 0.6, 0.17), (0.13, 0.6, 0.0), (-0.05, 0.6, 0.57)], 0)         # shortBox
    P += _box_from_top([(-0.53, 1.2, 0.09), (0.04, 1.2, -0.09), (-0.14, 1.2, -0.67), (-0.71, 1.2, -0.49)], 0)   # tallBox
    prims = np.concatenate(P)
    mats = make_materials([("lambert", WHITE), ("lambert", GREEN), ("lambert", RED)])
    light = _light((-0.24, 1.98, 0.16), (-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16), (10., 10., 10.))
    cam = PinholeCamera((0, 1, 3), (0, 0, -1), 120.0, width / height)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, integrator, name="cornell")


def _room(lo, hi, open_side="+z"):
    """5 inward-facing quads of the box [lo,hi]^3, open toward +z."""
    l, h = float(lo), float(hi)
    P = [quad_facing((l, l, l), (l, l, h), (h, l, h), (h, l, l), (0, 1, 0), 0),      # floor
         quad_facing((l, h, l), (l, h, h), (h, h, h), (h, h, l), (0, -1, 0), 0),     # ceiling
         quad_facing((l, l, l), (h, l, l), (h, h, l), (l, h, l), (0, 0, 1), 0),      # back wall
         quad_facing((l, l, l), (l, l, h), (l, h, h), (l, h, l), (1, 0, 0), 2),      # left (red)
         quad_facing((h, l, l), (h, l, h), (h, h, h), (h, h, l), (-1, 0, 0), 1)]     # right (green)
    return P


def c1_cube(width=640, height=480, max_depth=5):
    v, f = load_mesh("cube")
    P = [triangles_from_mesh(v, f, 0)] + _room(-3, 3)
    prims = np.concatenate(P)
    mats = make_materials([("lambert", WHITE), ("lambert", GREEN), ("lambert", RED)])
    light = _light((-1, 2.98, 1), (-1, 2.98, -1), (1, 2.98, -1), (1, 2.98, 1), (10., 10., 10.))
    cam = PinholeCamera((0, 0, 8), (0, 0, -1), 120.0, width / height)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, PATH_INTEGRATOR, name="c1_cube",
                     meta={"spp": 16})


def c2_spot(width=1920, height=1080, max_depth=5):
    v, f = load_mesh("spot")
    P = [triangles_from_mesh(v, f, 0),
         quad_facing((-4, -0.74, -3), (-4, -0.74, 3), (4, -0.74, 3), (4, -0.74, -3), (0, 1, 0), 0),       # floor
         quad_facing((-4, -0.74, 2.5), (4, -0.74, 2.5), (4, 4, 2.5), (-4, 4, 2.5), (0, 0, -1), 0)]       # back wall
    prims = np.concatenate(P)
    mats = make_materials([("lambert", WHITE)])
    light = _light((-0.5, 2.5, 0.5), (-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (10., 10., 10.))
    cam = PinholeCamera((0, 0.1, -2.6), (0, 0, 1), 120.0, 16.0 / 9.0)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, PATH_INTEGRATOR, name="c2_spot",
                     meta={"spp": 64})


def c3_renault(width=1920, height=1080, max_depth=5):
    v, f = load_mesh("renault")
    tri = triangles_from_mesh(v, f, 0)
    idx = np.arange(len(tri), dtype=np.uint64)
    tri["material"] = (((idx * np.uint64(2654435761)) >> np.uint64(16)) % np.uint64(3)).astype(np.int32)
    floor = quad_facing((-8, -0.25, -8), (-8, -0.25, 8), (8, -0.25, 8), (8, -0.25, -8), (0, 1, 0), 0)
    prims = np.concatenate([tri, floor])
    mats = make_materials([("lambert", (0.8, 0.8, 0.8)), ("metal", (0.7, 0.6, 0.5), 0.1), ("spectrans", (1., 1., 1.), 1.0, 1.5)])
    light = _light((-1, 4.0, 1), (-1, 4.0, -1), (1, 4.0, -1), (1, 4.0, 1), (10., 10., 10.))
    cam = PinholeCamera((4.5, 1.6, 5.5), (-0.62, -0.18, -0.76), 120.0, 16.0 / 9.0)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, NEW_PATH_TRACER, name="c3_renault",
                     meta={"spp": 256})


def c4_spheres(width=3840, height=2160, max_depth=5, grid=316, seed=42):
    """RandomScene recipe (RenderTest/Sample/RayTracing.fs:384-415) scaled to grid x grid cells:
    radius 0.2, centre (a+0.9xi, 0.2, b+xi); 80% Lambert(xi*xi) / 15% Metal(0.5(1+xi), fuzz 0.5xi)
    / 5% SpecularTransmission(1 -> 1.5); ground sphere r=1000.  numpy PCG64(seed) drives xi."""
    rng = np.random.default_rng(seed)
    half = grid // 2
    a, b = np.meshgrid(np.arange(-half, grid - half), np.arange(-half, grid - half), indexing="ij")
    n = a.size
    choose = rng.random(n)
    cx = a.reshape(-1) + 0.9 * rng.random(n)
    cz = b.reshape(-1) + rng.random(n)
    centers = np.stack([cx, np.full(n, 0.2), cz], 1)
    specs = []
    lam = rng.random((n, 3)) * rng.random((n, 3))
    met = 0.5 * (1.0 + rng.random((n, 3)))
    fuzz = 0.5 * rng.random(n)
    mats = np.zeros(n + 1, dtype=make_materials([]).dtype)
    kind = np.where(choose < 0.8, 0, np.where(choose < 0.95, 1, 2))
    mats["kind"][:n] = kind
    mats["albedo"][:n] = np.where((kind == 0)[:, None], lam, np.where((kind == 1)[:, None], met, 1.0))
    mats["fuzz"][:n] = np.where(kind == 1, fuzz, 0.0)
    mats["ei"][:n] = 1.0
    mats["et"][:n] = 1.5
    mats["kind"][n] = 0
    mats["albedo"][n] = (0.5, 0.5, 0.5)
    del specs
    small = sphere_prims(centers, 0.2, np.arange(n, dtype=np.int32))
    ground = sphere_prims([(0., -1000., 0.)], 1000.0, n)
    prims = np.concatenate([small, ground])
    light = _light((-25, 30, 25), (-25, 30, -25), (25, 30, -25), (25, 30, 25), (10., 10., 10.))
    pos = np.array([13., 2., 3.])
    cam = PinholeCamera(pos, -pos, 120.0, 16.0 / 9.0)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, NEW_PATH_TRACER, name="c4_spheres",
                     meta={"spp": 128})


def c5_soup(width=3840, height=2160, max_depth=5, instances=1708, cols=42, spacing=2.0, seed=7):
    """spot instanced `instances` times on a cols-wide grid with a random Y rotation (host-flattened:
    the reference has no live instancing, Shape/Box.fs:41-129 is commented out)."""
    v, f = load_mesh("spot")
    rng = np.random.default_rng(seed)
    nf = len(f)
    prims = make_prims(instances * nf + 1)
    rows = (instances + cols - 1) // cols
    for k in range(instances):
        th = rng.random() * 2.0 * np.pi
        c, s = np.cos(th), np.sin(th)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        off = np.array([(k % cols - (cols - 1) / 2.0) * spacing, 0.0, (k // cols - (rows - 1) / 2.0) * spacing])
        vv = v @ R.T + off
        prims[k * nf:(k + 1) * nf] = triangles_from_mesh(vv, f, 0)
    ext = cols * spacing
    prims[-1:] = quad_facing((-ext, -0.74, -ext), (-ext, -0.74, ext), (ext, -0.74, ext), (ext, -0.74, -ext), (0, 1, 0), 0)
    mats = make_materials([("lambert", WHITE)])
    light = _light((-20, 25, 20), (-20, 25, -20), (20, 25, -20), (20, 25, 20), (10., 10., 10.))
    cam = PinholeCamera((0, 14, -ext * 0.62), (0, -0.45, 1), 120.0, 16.0 / 9.0)
    return SceneDesc(prims, mats, light, cam, width, height, max_depth, PATH_INTEGRATOR, name="c5_soup",
                     meta={"spp": 64})


def perlin_tables(seed=7):
    """Perlin's static tables (RayTracing.fs:64-85): ranfloat = 256 uniforms, perm_x/y/z = Permute([0..255]) --
    the reference fills them from Random.Shared, here numpy PCG64(seed) with the same shuffle loop (:69-75)."""
    rng = np.random.default_rng(seed)
    ranfloat = rng.random(256)
    perm = np.zeros((3, 256), np.int32)
    for a in range(3):
        p = np.arange(256, dtype=np.int32)
        for i in range(255, -1, -1):
            targ = int(rng.random() * float(i + 1))
            p[i], p[targ] = p[targ], p[i]
        perm[a] = p
    return ranfloat, perm.reshape(-1)


def random_scene(width=400, height=200, max_depth=50, seed=42, ground="noise", aperture=0.0, cells=range(-1, 12)):
    """RandomScene (RenderTest/Sample/RayTracing.fs:384-415) behind the camera of DoRayTrace (:427-433): one small
    sphere per (a, b) cell -- 80 % Lambertian(ConstantTexture(xi*xi)), 15 % Metal(0.5(1+xi), fuzz 0.5 xi), 5 %
    Dielectric(1.5), skipped within 0.9 of (4, 0.2, 0) -- then the ground sphere r = 1000 (NoiseTexture; "checker" and
    "grey" are the two alternatives commented out at :410-411) and the three big spheres.  numpy PCG64(seed) stands
    in for `new System.Random()`, drawn in the reference's order."""
    rng = np.random.default_rng(seed)
    centers, radii, specs = [], [], []
    for a in cells:                                  # Array.allPairs [|-1..11|] [|-1..11|]: a-major
        for b in cells:
            choose_mat = rng.random()
            center = np.array([a + 0.9 * rng.random(), 0.2, b + rng.random()])
            if np.sqrt(((center - np.array([4, 0.2, 0])) ** 2).sum()) > 0.9:
                if choose_mat < 0.8:
                    specs.append(("lambert", (rng.random() * rng.random(), rng.random() * rng.random(), rng.random() * rng.random())))
                elif choose_mat < 0.95:
                    alb = (0.5 * (1. + rng.random()), 0.5 * (1. + rng.random()), 0.5 * (1. + rng.random()))
                    specs.append(("metal", alb, 0.5 * rng.random()))
                else:
                    specs.append(("dielectric", 1.5))
                centers.append(center)
                radii.append(0.2)
    specs.append({"noise": ("noise",), "checker": ("checker", (0.2, 0.3, 0.1), (0.9, 0.9, 0.9)),
                  "grey": ("lambert", (0.5, 0.5, 0.5))}[ground])
    centers.append((0., -1000., 0.)); radii.append(1000.)
    specs.append(("dielectric", 1.5)); centers.append((0., 1., 0.)); radii.append(1.0)
    specs.append(("lambert", (0.4, 0.2, 0.1))); centers.append((-4., 1., 0.)); radii.append(1.0)
    specs.append(("metal", (0.7, 0.6, 0.5), 0.0)); centers.append((4., 1., 0.)); radii.append(1.0)
    prims = sphere_prims(np.array(centers, float), np.array(radii, float), np.arange(len(specs), dtype=np.int32))
    mats = make_materials(specs)
    cam = RayTraceCamera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, width / height, aperture)
    rf, pm = perlin_tables(seed + 1)
    return SceneDesc(prims, mats, None, None, width, height, max_depth, SKY_TRACER, name="random_scene",
                     meta={"spp": 9}, sky=SkyTracer(cam, rf, pm))


WORKLOADS = {"cornell": cornell, "c1_cube": c1_cube, "c2_spot": c2_spot, "c3_renault": c3_renault,
             "c4_spheres": c4_spheres, "c5_soup": c5_soup, "random_scene": random_scene}
