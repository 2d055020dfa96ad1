"""A second, independent restatement of the reference integrator in pure Python (brute-force
intersection, no BVH), written from the F# sources again: used only to cross-check the C oracle
on tiny scenes (tests/test_oracle_kat.py).  Python floats are IEEE doubles and every expression
keeps the F# association, so agreement is expected to the last bit."""
import math

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = list(ctr)
    k0, k1 = key
    for r in range(10):
        if r:
            k0 = (k0 + W0) & MASK
            k1 = (k1 + W1) & MASK
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
    return c


def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def mul(v, s): return (v[0] * s, v[1] * s, v[2] * s)
def div(v, s): return (v[0] / s, v[1] / s, v[2] / s)
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, v): return (a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0])
def length(a): return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def normalize(a):
    l = length(a)
    return (0.0, 0.0, 0.0) if l == 0.0 else (a[0] / l, a[1] / l, a[2] / l)


def tri_hit(v0, v1, v2, o, d, tmin):                     # Trangle.fs:120-155
    e1, e2 = sub(v1, v0), sub(v2, v0)
    s1 = cross(d, e2)
    divisor = dot(s1, e1)
    if abs(divisor) < 1e-6:
        return None
    inv = 1. / divisor
    dd = sub(o, v0)
    b1 = dot(dd, s1) * inv
    if b1 < 0. or b1 > 1.:
        return None
    s2 = cross(dd, e1)
    b2 = dot(d, s2) * inv
    if b2 < 0. or (b1 + b2) >= 1.:
        return None
    t = dot(e2, s2) * inv
    return t if t > tmin else None


def tri_normal(v0, v1, v2):                              # Trangle.fs:108-112
    a = cross(sub(v1, v0), sub(v2, v0))
    return div(a, length(a))


class TinyScene:
    """Rect-only scene, brute-force closest hit (valid for rays in general position)."""

    def __init__(self, rects, albedos, light_p, light_n, light_color, cam12, width, height, max_depth):
        self.rects, self.albedos = rects, albedos
        self.lp, self.ln, self.lc = light_p, light_n, light_color
        self.cam = cam12
        self.w, self.h, self.D = width, height, max_depth
        p0, p1, p2, p3 = light_p
        self.area = length(cross(sub(p1, p0), sub(p2, p0))) * 0.5 + length(cross(sub(p2, p0), sub(p3, p0))) * 0.5

    def hit(self, o, d, tmin, tmax):
        best = None
        for i, (v0, v1, v2, v3, m) in enumerate(self.rects):
            t = tri_hit(v0, v1, v2, o, d, tmin)                  # Rect.fs:26-31
            n = tri_normal(v0, v1, v2)
            if t is None:
                t = tri_hit(v0, v2, v3, o, d, tmin)
                n = tri_normal(v0, v2, v3)
            if t is not None and t < tmax and (best is None or t < best[0]):
                best = (t, add(o, mul(d, t)), n, m, i)
        return best

    def uniforms(self, pix, sample, dim, it, seed):
        o = philox4x32_10([pix, sample, dim, it], [seed & MASK, seed >> 32])
        return [x * (1.0 / 4294967296.0) for x in o]

    def random_in_unit_sphere(self, nm, pix, sample, dim, seed):     # Material.fs:9-14
        p = (20., 20., 20.)
        it = 0
        while dot(p, p) >= 1.0 or dot(nm, p) <= 0.:
            u = self.uniforms(pix, sample, dim, it, seed)
            it += 1
            p = sub(mul((u[0], u[1], u[2]), 2.0), (1., 1., 1.))
        return p

    def sample_light(self, pix, sample, k, seed):                    # Rect.fs:33-38, Trangle.fs:157-169
        u = self.uniforms(pix, sample, 2 + 2 * k, 0, seed)
        p0, p1, p2, p3 = self.lp
        v0, v1, v2 = (p0, p1, p2) if u[0] < 0.5 else (p0, p2, p3)
        tu, tv = u[1], u[2]
        uu, vv = (1. - tu, 1. - tv) if tu + tv > 1. else (tu, tv)
        sq = math.sqrt(1. - uu)
        return add(add(v0, mul(sub(v1, v0), 1. - sq)), mul(sub(v2, v0), vv * sq))

    def trace(self, o, d, depth, pix, sample, seed):                 # Integrators.fs:107-138
        h = self.hit(o, d, 1e-6, 99999999.)
        if h is None or depth < 0:
            return (0., 0., 0.)
        t, point, n, m, _ = h
        k = self.D - depth
        a = self.albedos[m]
        wi = normalize(self.random_in_unit_sphere(n, pix, sample, 1 + 2 * k, seed))
        ei = dot(n, wi)
        INVPI, TWOPI = 1. / math.pi, 2. * math.pi
        col = tuple(TWOPI * (ei * (INVPI * c)) for c in a)
        lp = self.sample_light(pix, sample, k, seed)
        toLight = sub(lp, point)
        dist = length(toLight)
        unit = div(toLight, dist)
        sh = self.hit(point, unit, 1e-6, dist - 1e-6)
        if sh is not None:
            l = (0., 0., 0.)
        else:
            cos_o = dot(toLight, self.ln)
            if cos_o < 0.:
                solid = abs(cos_o) * self.area / dot(toLight, toLight)
                L = tuple(solid * c for c in self.lc)
            else:
                L = (0., 0., 0.)
            dn = dot(unit, n)
            l = tuple(dn * c for c in L)
        pdf_li = 1. / self.area
        li = self.trace(point, wi, depth - 1, pix, sample, seed)
        return tuple(((l[c] / pdf_li + li[c]) * col[c]) / 1. for c in range(3))

    def trace_path(self, i, j, sample, seed):
        pix = j * self.w + i
        u4 = self.uniforms(pix, sample, 0, 0, seed)
        u = (float(i) + u4[0]) / float(self.w)
        v = (float(j) + u4[1]) / float(self.h)
        pos, tl, right, down = self.cam[0:3], self.cam[3:6], self.cam[6:9], self.cam[9:12]
        target = add(add(tl, mul(right, u)), mul(down, v))
        d = normalize(sub(target, pos))
        return self.trace(tuple(pos), d, self.D, pix, sample, seed)
