"""ctypes binding of the CPU ORACLE (oracle/mafrix_oracle.c).  TEST INFRASTRUCTURE ONLY --
PARITY UNPINNED (see mafrix_oracle.h): importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmafrix_oracle.so")

PRIM_DTYPE = np.dtype([("kind", "<i4"), ("material", "<i4"), ("v", "<f8", (12,))])
MATERIAL_DTYPE = np.dtype([("kind", "<i4"), ("pad", "<i4"), ("albedo", "<f8", (3,)),
                           ("fuzz", "<f8"), ("ei", "<f8"), ("et", "<f8")])
NODE_DTYPE = np.dtype([("pmin", "<f8", (3,)), ("pmax", "<f8", (3,)), ("first", "<i4"), ("count", "<i4")])


class OrcStats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("wasted_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("ref_nodes", C.c_uint64), ("ref_prims", C.c_uint64),
                ("ord_rays", C.c_uint64 * 2), ("ord_nodes", C.c_uint64 * 2),
                ("ord_tris", C.c_uint64 * 2), ("ord_spheres", C.c_uint64 * 2)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


def build(force=False):
    srcs = ["mafrix_oracle.c", "mafrix_oracle_sky.c", "mafrix_oracle.h"]
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.orc_scene_create.restype = P
        L.orc_scene_create.argtypes = [P, C.c_int, P, C.c_int, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_scene_destroy.argtypes = [P]
        L.orc_scene_node_slots.argtypes = [P]
        L.orc_scene_get_bvh.argtypes = [P, P, P]
        L.orc_scene_set_bvh.argtypes = [P, P, P]
        L.orc_camera_pinhole.argtypes = [P, P, C.c_double, C.c_double, P]
        L.orc_bvh_hit.argtypes = [P, C.c_int, P, P, C.c_double, C.c_double, P, P, P, P, P]
        L.orc_trace_primary.argtypes = [P, C.c_int, P, P, P]
        L.orc_sample.argtypes = [P, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 P, P, C.c_int]
        L.orc_trace_path.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_uint64, P]
        L.orc_film_add_sample.argtypes = [P, P, P, C.c_int, C.c_double]
        L.orc_tonemap_rgba8.argtypes = [P, C.c_int, C.c_int, P]
        L.orc_philox4x32_10.argtypes = [P, P, P]
        L.orc_max_threads.restype = C.c_int
        L.orc_camera_lens.argtypes = [P, P, P, C.c_double, C.c_double, C.c_double, C.c_double, P]
        L.orc_sky_create.restype = P
        L.orc_sky_create.argtypes = [P, C.c_int, P, C.c_int, P, P, P, C.c_int, C.c_int, C.c_int]
        L.orc_sky_destroy.argtypes = [P]
        L.orc_sky_list_hit.argtypes = [P, C.c_int, P, P, C.c_double, C.c_double, P, P]
        L.orc_sky_trace_primary.argtypes = [P, C.c_int, P, P, P]
        L.orc_sky_trace_path.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_uint64, P, P]
        L.orc_sky_sample.argtypes = [P, C.c_int, C.c_uint64, C.c_int, C.c_int, P, P]
        L.orc_sky_display_rgba8.argtypes = [P, C.c_int, C.c_int, P]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def camera_pinhole(pos, dir, fov, aspect):
    out = np.zeros(12)
    pos = np.ascontiguousarray(pos, np.float64)
    dir = np.ascontiguousarray(dir, np.float64)
    lib().orc_camera_pinhole(_p(pos), _p(dir), float(fov), float(aspect), _p(out))
    return out


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def max_threads():
    return lib().orc_max_threads()


class OracleScene:
    """The oracle's `Scene`: built from any object with the SceneDesc attributes
    (prims, materials, light.{p,normal,color}, camera.{pos_arg,dir_arg,fov,aspect}, width, height,
    max_depth, integrator).  The camera is re-derived by the oracle's own PinholeCamera restatement."""

    def __init__(self, desc, cam12=None):
        self.desc = desc
        self.width, self.height = int(desc.width), int(desc.height)
        self.prims = np.ascontiguousarray(desc.prims).view(PRIM_DTYPE) if desc.prims.dtype.itemsize == 104 else None
        self.mats = np.ascontiguousarray(desc.materials).view(MATERIAL_DTYPE)
        assert self.prims is not None
        self.cam = camera_pinhole(desc.camera.pos_arg, desc.camera.dir_arg, desc.camera.fov, desc.camera.aspect) \
            if cam12 is None else np.ascontiguousarray(cam12, np.float64)
        lp = np.ascontiguousarray(desc.light.p, np.float64).reshape(12)
        ln = np.ascontiguousarray(desc.light.normal, np.float64)
        lc = np.ascontiguousarray(desc.light.color, np.float64)
        self._h = lib().orc_scene_create(_p(self.prims), len(self.prims), _p(self.mats), len(self.mats),
                                         _p(lp), _p(ln), _p(lc), _p(self.cam), self.width, self.height,
                                         int(desc.max_depth), int(desc.integrator))

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh(self):
        n = lib().orc_scene_node_slots(self._h)
        nodes = np.zeros(n, NODE_DTYPE)
        idx = np.zeros(len(self.prims), np.int32)
        lib().orc_scene_get_bvh(self._h, _p(nodes), _p(idx))
        return nodes, idx

    def set_bvh(self, nodes, indices):
        nodes = np.ascontiguousarray(nodes).view(NODE_DTYPE)
        indices = np.ascontiguousarray(indices, np.int32)
        lib().orc_scene_set_bvh(self._h, _p(nodes), _p(indices))

    def hit(self, origins, dirs, tmin, tmax, want_geometry=False):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        n = len(o)
        prim = np.zeros(n, np.int32)
        sub = np.zeros(n, np.int32)
        t = np.zeros(n)
        pt = np.zeros((n, 3)) if want_geometry else None
        nm = np.zeros((n, 3)) if want_geometry else None
        lib().orc_bvh_hit(self._h, n, _p(o), _p(d), float(tmin), float(tmax), _p(prim), _p(sub), _p(t), _p(pt), _p(nm))
        return (prim, sub, t, pt, nm) if want_geometry else (prim, sub, t)

    def trace_primary(self, uv=None):
        if uv is None:
            n = self.width * self.height
        else:
            uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
            n = len(uv)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n)
        lib().orc_trace_primary(self._h, n, _p(uv), _p(prim), _p(t))
        return prim, t

    def sample(self, n, seed=1, first_sample=0, region=None, threads=0, stats=False, count_ordered=False, out=None):
        """PixelIntegrator.Sample(n) -> Color[w,h] as (width, height, 4) f64."""
        tex = np.zeros((self.width, self.height, 4)) if out is None else out
        x0, y0, x1, y1 = region if region is not None else (0, 0, self.width, self.height)
        st = OrcStats() if (stats or count_ordered) else None
        lib().orc_sample(self._h, int(n), int(seed), int(first_sample), x0, y0, x1, y1, int(threads),
                         _p(tex), C.byref(st) if st is not None else None, int(count_ordered))
        return (tex, st.as_dict()) if st is not None else tex

    def trace_path(self, px, py, sample, seed=1):
        rgb = np.zeros(3)
        lib().orc_trace_path(self._h, int(px), int(py), int(sample), int(seed), _p(rgb))
        return rgb


def camera_lens(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist):
    """RayTraceCamera constructor (RayTracing.fs:335-358) -> 19 doubles: origin, lower_left, horizontal, vertical,
    u, v, lens_radius."""
    out = np.zeros(19)
    a = [np.ascontiguousarray(x, np.float64) for x in (lookfrom, lookat, vup)]
    lib().orc_camera_lens(_p(a[0]), _p(a[1]), _p(a[2]), float(vfov), float(aspect), float(aperture), float(focus_dist), _p(out))
    return out


class OracleSkyScene:
    """The sphere sample (RenderTest/Sample/RayTracing.fs): built from a SceneDesc whose integrator is the sky
    tracer -- prims (spheres), materials, sky.camera (a RayTraceCamera mirror with its constructor arguments),
    sky.ranfloat / sky.perm, width, height, max_depth.  The camera is re-derived by the oracle's own restatement."""

    def __init__(self, desc, cam19=None):
        self.desc = desc
        self.width, self.height = int(desc.width), int(desc.height)
        self.prims = np.ascontiguousarray(desc.prims).view(PRIM_DTYPE)
        self.mats = np.ascontiguousarray(desc.materials).view(MATERIAL_DTYPE)
        c = desc.sky.camera
        self.cam = camera_lens(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect, c.aperture, c.focus_dist) \
            if cam19 is None else np.ascontiguousarray(cam19, np.float64)
        rf = None if desc.sky.ranfloat is None else np.ascontiguousarray(desc.sky.ranfloat, np.float64)
        pm = None if desc.sky.perm is None else np.ascontiguousarray(desc.sky.perm, np.int32)
        self._h = lib().orc_sky_create(_p(self.prims), len(self.prims), _p(self.mats), len(self.mats), _p(self.cam),
                                       _p(rf), _p(pm), self.width, self.height, int(desc.max_depth))

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_sky_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def hit(self, origins, dirs, tmin=0.00001, tmax=10000000.):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o))
        lib().orc_sky_list_hit(self._h, len(o), _p(o), _p(d), float(tmin), float(tmax), _p(prim), _p(t))
        return prim, t

    def trace_primary(self, uv=None):
        if uv is None:
            n = self.width * self.height
        else:
            uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
            n = len(uv)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n)
        lib().orc_sky_trace_primary(self._h, n, _p(uv), _p(prim), _p(t))
        return prim, t

    def sample(self, n, seed=1, first_sample=0, threads=0, stats=False):
        tex = np.zeros((self.width, self.height, 4))
        rays = C.c_uint64()
        lib().orc_sky_sample(self._h, int(n), int(seed), int(first_sample), int(threads), _p(tex), C.byref(rays))
        return (tex, {"closest_rays": rays.value}) if stats else tex

    def trace_path(self, px, py, sample, seed=1):
        rgb = np.zeros(3)
        rays = C.c_uint64()
        lib().orc_sky_trace_path(self._h, int(px), int(py), int(sample), int(seed), _p(rgb), C.byref(rays))
        return rgb


def film_add_sample(sum_, frame, frame_count):
    target = np.zeros_like(sum_)
    lib().orc_film_add_sample(_p(sum_), _p(np.ascontiguousarray(frame)), _p(target), sum_.size // 4, float(frame_count))
    return target


def sky_display_rgba8(texture_wh):
    """sqrt + int(255.99 c) + vertical flip of the sphere sample (RayTracing.fs:456-460) -> (height, width, 4) uint8."""
    w, h = texture_wh.shape[0], texture_wh.shape[1]
    out = np.zeros((h, w, 4), np.uint8)
    lib().orc_sky_display_rgba8(_p(np.ascontiguousarray(texture_wh)), w, h, _p(out))
    return out


def tonemap_rgba8(texture_wh):
    w, h = texture_wh.shape[0], texture_wh.shape[1]
    out = np.zeros((h, w, 4), np.uint8)
    lib().orc_tonemap_rgba8(_p(np.ascontiguousarray(texture_wh)), w, h, _p(out))
    return out
