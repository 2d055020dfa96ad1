"""Second, independent restatement of the sphere sample's integrator (GetColor and everything under it,
/root/reference/RenderTest/Sample/RayTracing.fs) in pure Python, written from the F# again and structured like the
F# (closures per material, a recursive GetColor): cross-checks oracle/mafrix_oracle_sky.c on tiny scenes.  Python
floats are IEEE doubles and every expression keeps the F# association, so agreement is expected to the last bit.
The random stream is the repo's Philox layout (DESIGN.md "RNG"), and Math.Pow(x, 5) is the product x2*x2*x like
in the oracle (its header says why)."""
import math

from .pyref import add, cross, div, dot, mul, normalize, philox4x32_10, sub

CAP = 128


def uniforms(pixel, sample, dim, it, seed):
    o = philox4x32_10((pixel, sample, dim, it), (seed & 0xFFFFFFFF, seed >> 32))
    return [x * (1.0 / 4294967296.0) for x in o]


class Ray:                                                   # :14-21
    def __init__(self, origin, direc):
        self.A, self.B = origin, normalize(direc)

    def at(self, t):
        return add(self.A, mul(self.B, t))


def lens_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist):      # :335-358
    theta = vfov * math.pi / 180.
    half_height = math.tan(theta / 2.)
    half_width = aspect * half_height
    w = normalize(sub(lookfrom, lookat))
    u = normalize(cross(vup, w))
    v = normalize(cross(w, u))
    p = sub(sub(sub(lookfrom, mul(u, focus_dist * half_width)), mul(v, focus_dist * half_height)), mul(w, focus_dist))
    return dict(origin=tuple(lookfrom), lower_left=p, horizontal=mul(u, 2. * focus_dist * half_width),
                vertical=mul(v, 2. * focus_dist * half_height), u=u, v=v, lens_radius=aperture / 2.0)


def sphere_hit(center, radius, r, tmin, tmax):              # :188-207
    oc = sub(r.A, center)
    a = dot(r.B, r.B)
    b = 2.0 * dot(oc, r.B)
    c = dot(oc, oc) - radius * radius
    disc = b * b - 4.0 * a * c
    if disc > 0:
        for tmp in ((-b - math.sqrt(disc)) / (2.0 * a), (-b + math.sqrt(disc)) / (2.0 * a)):
            if tmp < tmax and tmp > tmin:
                p = r.at(tmp)
                return tmp, p, div(sub(p, center), radius)
    return None


def list_hit(spheres, r, tmin, tmax):                        # :256-258, Array.minBy keeps the first minimum
    best, best_key = None, None
    for i, (c, rad, m) in enumerate(spheres):
        h = sphere_hit(c, rad, r, tmin, tmax)
        key = h[0] if h else tmax
        if best_key is None or key < best_key:
            best_key, best = key, (i, h, m) if h else None
    return best


def unit_ball(rng, dim):                                     # :261-266
    p = (20., 20., 20.)
    it = 0
    while dot(p, p) >= 1.0:
        if it >= CAP:
            return (0., 0., 0.)
        u = rng(dim, it)
        it += 1
        p = sub(mul((u[0], u[1], u[2]), 2.0), (1., 1., 1.))
    return p


def unit_disk(rng):                                          # :327-333
    p, d, it = (0., 0., 0.), 1.0, 0
    while d >= 1.0:
        if it >= CAP:
            return (0., 0., 0.)
        u = rng(0, 1 + it)
        it += 1
        p = sub(mul((u[0], u[1], 0.), 2.0), (1., 1., 0.))
        d = dot(p, p)
    return p


def reflect(v, n):                                           # :268
    return sub(v, mul(n, 2.0 * dot(v, n)))


def texture(mat, tables, p):                                 # :50-61, :86-99
    kind = mat["kind"]
    if kind == "checker":
        sines = math.sin(10. * p[0]) * math.sin(10. * p[1]) * math.sin(10. * p[2])
        return mat["odd"] if sines < 0. else mat["even"]
    if kind == "noise":
        rf, pm = tables
        i, j, k = int(4. * p[0]) & 255, int(4. * p[1]) & 255, int(4. * p[2]) & 255
        nz = float(rf[int(pm[i]) ^ int(pm[256 + j]) ^ int(pm[512 + k])])
        return (nz, nz, nz)
    return mat["albedo"]


def scatter(mat, tables, ray, t, p, normal, rng, k):
    kind = mat["kind"]
    if kind == "metal":                                      # :291-299
        fuzz = mat["fuzz"] if mat["fuzz"] < 1.0 else 1.0
        reflected = reflect(normalize(ray.B), normal)
        sc = Ray(p, add(reflected, mul(unit_ball(rng, 1 + 2 * k), fuzz)))
        return dot(sc.B, normal) > 0, mat["albedo"], sc
    if kind == "dielectric":                                 # :300-325
        ri = mat["ri"]
        reflected = reflect(ray.B, normal)
        if dot(ray.B, normal) > 0:
            outward, nint, cosine = (-normal[0], -normal[1], -normal[2]), ri, ri * dot(ray.B, normal)
        else:
            outward, nint, cosine = normal, 1.0 / ri, -dot(ray.B, normal)
        uv = normalize(ray.B)                                # Refract, :269-276
        dt = dot(uv, outward)
        disc = 1.0 - nint * nint * (1.0 - dt * dt)
        if disc > 0:
            ref_dir = sub(mul(sub(ray.B, mul(outward, dt)), nint), mul(outward, math.sqrt(disc)))
            r0 = (1. - ri) / (1. + ri)
            r1 = r0 * r0
            x = 1. - cosine
            x2 = x * x
            prob = r1 + (1. - r1) * ((x2 * x2) * x)          # Schlick, :277-280
        else:
            ref_dir, prob = (0., 0., 0.), 1.0
        coin = rng(2 + 2 * k, 0)[0]
        return True, (1., 1., 1.), Ray(p, reflected if coin < prob else ref_dir)
    target = add(normalize(normal), unit_ball(rng, 1 + 2 * k))      # Lambertian, :282-290
    return True, texture(mat, tables, p), Ray(p, target)


def get_color(spheres, tables, ray, depth, max_depth, rng):   # :367-382
    h = list_hit(spheres, ray, 0.00001, 10000000.)
    if h:
        _, (t, p, normal), mat = h
        ok, att, sc = scatter(mat, tables, ray, t, p, normal, rng, depth)
        if depth < max_depth and ok:
            c = get_color(spheres, tables, sc, depth + 1, max_depth, rng)
            return (c[0] * att[0], c[1] * att[1], c[2] * att[2])
        return (0., 0., 0.)
    unit = normalize(ray.B)
    t = 0.5 * (unit[1] + 1.0)
    return add(mul((1., 1., 1.), 1.0 - t), mul((0.5, 0.7, 1.0), t))


def get_ray(cam, s, t, rng):                                  # :360-364
    rd = mul(unit_disk(rng), cam["lens_radius"])
    offset = add(mul(cam["u"], rd[0]), mul(cam["v"], rd[1]))
    d = sub(sub(add(add(cam["lower_left"], mul(cam["horizontal"], s)), mul(cam["vertical"], t)), cam["origin"]), offset)
    return Ray(add(cam["origin"], offset), d)


def trace_path(spheres, tables, cam, width, height, max_depth, px, py, sample, seed):
    pixel = py * width + px

    def rng(dim, it):
        return uniforms(pixel, sample, dim, it, seed)
    u4 = rng(0, 0)
    u = (float(px) + u4[0]) / float(width)
    v = (float(py) + u4[1]) / float(height)
    return get_color(spheres, tables, get_ray(cam, u, v, rng), 0, max_depth, rng)
