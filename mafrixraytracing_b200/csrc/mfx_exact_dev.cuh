// mfx_exact_dev.cuh -- f64 device functions in the operation order of the F# reference.
// ONLY for translation units compiled with --fmad=false (mfx_exact.cu, mfx_hybrid.cu): a contracted a*b+c
// would round differently from RyuJIT's separate multiply and add.
#pragma once
#include "mfx_device.cuh"

typedef V3<double> D3;

struct HitX {
    int    slot;   // leaf-order primitive slot or -1
    int    sub;    // which triangle of a Rect
    double t;
};

// AABB.hit (Interfaces/IHitable.fs:18-54): divides by dir, `>= 0.` select (true for -0.0).
__device__ __forceinline__ bool aabb_hit_x(const NodeX &n, D3 o, D3 dir, double tMin, double tMax, double &entry)
{
    double tmin, tmax, tymin, tymax, tzmin, tzmax;
    if (dir.x >= 0.) { tmin = (n.pmin[0] - o.x) / dir.x; tmax = (n.pmax[0] - o.x) / dir.x; }
    else             { tmin = (n.pmax[0] - o.x) / dir.x; tmax = (n.pmin[0] - o.x) / dir.x; }
    if (dir.y >= 0.) { tymin = (n.pmin[1] - o.y) / dir.y; tymax = (n.pmax[1] - o.y) / dir.y; }
    else             { tymin = (n.pmax[1] - o.y) / dir.y; tymax = (n.pmin[1] - o.y) / dir.y; }
    if (tmin > tymax || tymin > tmax) return false;
    tmin = (tymin > tmin) ? tymin : tmin;
    tmax = (tymax < tmax) ? tymax : tmax;
    if (dir.z >= 0.) { tzmin = (n.pmin[2] - o.z) / dir.z; tzmax = (n.pmax[2] - o.z) / dir.z; }
    else             { tzmin = (n.pmax[2] - o.z) / dir.z; tzmax = (n.pmin[2] - o.z) / dir.z; }
    if (tmin > tzmax || tzmin > tmax) return false;
    tmin = (tzmin > tmin) ? tzmin : tmin;
    tmax = (tzmax < tmax) ? tzmax : tmax;
    entry = tmin;
    return tmin < tMax && tmax > tMin;
}

// Triangle.PreCalcu + Hit (Shape/Trangle.fs:120-155); tMax ignored (quirk Q2).
__device__ __forceinline__ bool tri_hit_x(D3 v0, D3 e1, D3 e2, D3 o, D3 dir, double tMin, double &t_out)
{
    const D3 s1 = cross(dir, e2);
    const double divisor = dot(s1, e1);
    if (fabs(divisor) < 1e-6) return false;
    const double inv = 1. / divisor;
    const D3 d = o - v0;
    const double b1 = dot(d, s1) * inv;
    if (b1 < 0. || b1 > 1.) return false;
    const D3 s2 = cross(d, e1);
    const double b2 = dot(dir, s2) * inv;
    if (b2 < 0. || (b1 + b2) >= 1.) return false;
    const double t = dot(e2, s2) * inv;
    if (t > tMin) { t_out = t; return true; }
    return false;
}

// Sphere.Hit (Shape/Sphere.fs:21-43)
__device__ __forceinline__ bool sphere_hit_x(D3 center, double radius, D3 o, D3 dir, double tMin, double tMax, double &t_out)
{
    const D3 oc = o - center;
    const double a = 1.;
    const double b = 2.0 * dot(oc, dir);
    const double c = dot(oc, oc) - radius * radius;
    const double disc = b * b - 4.0 * a * c;
    if (disc > 0) {
        const double root = sqrt(disc);
        const double q = (b < 0.) ? -0.5 * (b - root) : -0.5 * (b + root);
        const double t0 = q, t1 = c / q;
        const double lo = (t0 < t1) ? t0 : t1, hi = (t0 > t1) ? t0 : t1;
        if (lo >= tMin && lo < tMax) { t_out = lo; return true; }
        else if (hi > tMin && hi < tMax) { t_out = hi; return true; }
    }
    return false;
}

// IHitable.Hit on one leaf-order slot; Rect.Hit: tri1 if it hits ELSE tri2 (Rect.fs:26-31, quirk Q3).
__device__ __forceinline__ bool prim_hit_x(const PrimX &p, D3 o, D3 dir, double tMin, double tMax, double &t, int &sub)
{
    sub = 0;
    if (p.kind == 0) return tri_hit_x(ld3(p.v0), ld3(p.e1), ld3(p.e2), o, dir, tMin, t);
    if (p.kind == 1) {
        if (tri_hit_x(ld3(p.v0), ld3(p.e1), ld3(p.e2), o, dir, tMin, t)) return true;
        sub = 1;
        return tri_hit_x(ld3(p.v0), ld3(p.e2), ld3(p.e3), o, dir, tMin, t);
    }
    return sphere_hit_x(ld3(p.v0), p.e1[0], o, dir, tMin, tMax, t);
}

// Bvh.Hit (BvhNode.fs:62-83).  ANY: shadow query -- returns as soon as a leaf yields a hit record.
// COUNT: add visited node/primitive records to ctr (instrumented runs).
template <bool ANY, bool COUNT>
static __device__ HitX bvh_hit_x(const SceneX &sc, D3 o, D3 dir, double tMin, double tMax, unsigned long long *ctr)
{
    HitX best; best.slot = -1; best.sub = 0; best.t = 0.;
    int bestFirst = -1;
    int stack[40];
    int sp = 0;
    double e;
    NodeX node = sc.nodes[0];
    if (COUNT) ctr[0]++;
    if (!aabb_hit_x(node, o, dir, tMin, tMax, e)) return best;
    int cur = 0;
    for (;;) {
        if (node.count > MFX_LEAF_NODE_COUNT) {
            const int li = 2 * cur + 1, ri = 2 * cur + 2;
            const NodeX L = sc.nodes[li];
            const NodeX R = sc.nodes[ri];
            if (COUNT) ctr[0] += 2;
            double el, er;
            bool hl = aabb_hit_x(L, o, dir, tMin, tMax, el);
            bool hr = aabb_hit_x(R, o, dir, tMin, tMax, er);
            if (!ANY && best.slot >= 0) {
                // conservative t-shrink: a box whose entry is beyond the best hit by more than
                // rounding noise cannot hold a hit with t <= best.t (ties must survive, quirk Q1)
                const double lim = best.t + best.t * 1e-9;
                if (el > lim) hl = false;
                if (er > lim) hr = false;
            }
            if (hl && hr) {
                const bool rightNear = er < el;
                stack[sp++] = rightNear ? li : ri;
                cur = rightNear ? ri : li;
                node = rightNear ? R : L;
                continue;
            } else if (hl) { cur = li; node = L; continue; }
            else if (hr) { cur = ri; node = R; continue; }
        } else {
            // leaf: Array.map Hit |> Array.minBy (hit ? t : tMax) -- FIRST minimal key (BvhNode.fs:76-80)
            bool have = false, recHit = false; double bestKey = 0., recT = 0.; int recSlot = -1, recSub = 0;
            for (int k = 0; k < node.count; k++) {
                const PrimX p = sc.prims[node.first + k];
                if (COUNT) { if (p.kind == 2) ctr[2]++; else ctr[1] += (p.kind == 1) ? 2 : 1; }
                double t; int sub;
                const bool h = prim_hit_x(p, o, dir, tMin, tMax, t, sub);
                const double key = h ? t : tMax;
                if (!have || key < bestKey) { have = true; bestKey = key; recHit = h; recT = t; recSlot = node.first + k; recSub = sub; }
            }
            if (recHit) {
                if (ANY) { best.slot = recSlot; best.sub = recSub; best.t = recT; return best; }
                // interior combine `if l.t < r.t then l else r` (BvhNode.fs:69-70): on equal t the
                // record later in depth-first order wins == the leaf with the larger `first`
                if (best.slot < 0 || recT < best.t || (recT == best.t && node.first > bestFirst)) {
                    best.slot = recSlot; best.sub = recSub; best.t = recT; bestFirst = node.first;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
        node = sc.nodes[cur];
        if (COUNT) ctr[0]++;   // re-fetch of a deferred node (its box was already tested)
    }
    return best;
}

// ---------------------------------------------------------------- MFX_SKY_TRACER closest hit
// Sphere.Hit of the sphere sample (RenderTest/Sample/RayTracing.fs:188-207): a = d.d, near root first, strict bounds.
__device__ __forceinline__ bool sphere_hit_sky_x(D3 center, double radius, D3 o, D3 dir, double tMin, double tMax, double &t_out)
{
    const D3 oc = o - center;
    const double a = dot(dir, dir);
    const double b = 2.0 * dot(oc, dir);
    const double c = dot(oc, oc) - radius * radius;
    const double disc = b * b - 4.0 * a * c;
    if (disc > 0) {
        double tmp = (-b - sqrt(disc)) / (2.0 * a);
        if (tmp < tMax && tmp > tMin) { t_out = tmp; return true; }
        tmp = (-b + sqrt(disc)) / (2.0 * a);
        if (tmp < tMax && tmp > tMin) { t_out = tmp; return true; }
    }
    return false;
}

// ListHit (RayTracing.fs:256-258) tests EVERY sphere and keeps the first minimal t in list order.  The answer --
// smallest t, ties to the smaller list index -- does not depend on the order the spheres are visited in, so the
// kernel walks the scene's tree near-first with the conservative t-shrink of bvh_hit_x and applies the tie rule
// explicitly.  The tree's boxes were padded at flatten time (flatten_exact): the reference tests no boxes here, so
// a box must never reject a ray the sphere formula accepts.
template <bool COUNT>
static __device__ HitX sky_hit_x(const SceneX &sc, D3 o, D3 dir, double tMin, double tMax, unsigned long long *ctr)
{
    HitX best; best.slot = -1; best.sub = 0; best.t = 0.;
    int bestRef = 0x7fffffff;
    int stack[40];
    int sp = 0;
    double e;
    NodeX node = sc.nodes[0];
    if (COUNT) ctr[0]++;
    if (!aabb_hit_x(node, o, dir, tMin, tMax, e)) return best;
    int cur = 0;
    for (;;) {
        if (node.count > MFX_LEAF_NODE_COUNT) {
            const int li = 2 * cur + 1, ri = 2 * cur + 2;
            const NodeX L = sc.nodes[li];
            const NodeX R = sc.nodes[ri];
            if (COUNT) ctr[0] += 2;
            double el, er;
            bool hl = aabb_hit_x(L, o, dir, tMin, tMax, el);
            bool hr = aabb_hit_x(R, o, dir, tMin, tMax, er);
            if (best.slot >= 0) {
                const double lim = best.t + best.t * 1e-9;
                if (el > lim) hl = false;
                if (er > lim) hr = false;
            }
            if (hl && hr) {
                const bool rightNear = er < el;
                stack[sp++] = rightNear ? li : ri;
                cur = rightNear ? ri : li;
                node = rightNear ? R : L;
                continue;
            } else if (hl) { cur = li; node = L; continue; }
            else if (hr) { cur = ri; node = R; continue; }
        } else {
            for (int k = 0; k < node.count; k++) {
                const PrimX p = sc.prims[node.first + k];
                if (COUNT) ctr[2]++;
                double t;
                if (sphere_hit_sky_x(ld3(p.v0), p.e1[0], o, dir, tMin, tMax, t)) {
                    const int ref = sc.ref_id[node.first + k];
                    if (best.slot < 0 || t < best.t || (t == best.t && ref < bestRef)) { best.slot = node.first + k; best.t = t; bestRef = ref; }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
        node = sc.nodes[cur];
        if (COUNT) ctr[0]++;
    }
    return best;
}

// ---------------------------------------------------------------- RNG-driven samplers
struct RngX { uint32_t pixel, sample, k0, k1; };

__device__ __forceinline__ void rng_draw_x(const RngX &g, uint32_t dim, uint32_t iter, double (&u)[4])
{
    uint32_t o[4];
    philox4x32_10(g.pixel, g.sample, dim, iter, g.k0, g.k1, o);
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = u32_to_unit_f64(o[i]);
}

// GetRandomInUnitSphere (Materials/Material.fs:9-14), capped like the oracle.
static __device__ D3 random_in_unit_sphere_x(D3 nm, const RngX &g, uint32_t dim)
{
    D3 p = mk3<double>(20., 20., 20.);
    uint32_t it = 0;
    while (dot(p, p) >= 1.0 || dot(nm, p) <= 0.) {
        if (it >= MFX_REJECTION_CAP) return nm;
        double u[4];
        rng_draw_x(g, dim, it++, u);
        p = mk3<double>(u[0], u[1], u[2]) * 2.0 - mk3<double>(1., 1., 1.);
    }
    return p;
}

__device__ __forceinline__ D3 normalize_x(D3 a)     // Vector.Normalize, Point.fs:52-56
{
    const double l = sqrt(len2(a));
    if (l == 0.0) return mk3<double>(0., 0., 0.);
    return mk3<double>(a.x / l, a.y / l, a.z / l);
}

__device__ __forceinline__ D3 reflect_x(D3 v, D3 n) { return v - n * (2.0 * dot(v, n)); }   // Material.fs:16

__device__ __forceinline__ double fmax_fs(double a, double b) { return a > b ? a : (b > a ? b : (a != a ? a : b)); }

// FresnelDielectric.Evaluate (Material.fs:74-96)
static __device__ double fresnel_x(double eta_i, double eta_t, double cosi)
{
    double ei, et;
    if (cosi > 0.) { ei = eta_i; et = eta_t; } else { ei = eta_t; et = eta_i; }
    const double sint = ei / et * sqrt(fmax_fs(0., 1. - cosi * cosi));
    if (sint >= 1.) return 1.0;
    const double cost = sqrt(fmax_fs(0., 1. - sint * sint));
    const double ci = fabs(cosi);
    const double rparl = ((et * ci) - (ei * cost)) / ((et * ci) + (ei * cost));
    const double rperp = ((ei * ci) - (et * cost)) / ((ei * ci) + (et * cost));
    return (rparl * rparl + rperp * rperp) / 2.;
}

// Triangle.SamplePoint (Trangle.fs:157-169)
__device__ __forceinline__ D3 tri_sample_x(const TriSampleX &t, double tu, double tv)
{
    double u, v;
    if (tu + tv > 1.) { u = 1. - tu; v = 1. - tv; } else { u = tu; v = tv; }
    const double sq = sqrt(1. - u);
    const double s1 = 1. - sq;
    const double s2 = v * sq;
    return (ld3(t.v0) + ld3(t.e1) * s1) + ld3(t.e2) * s2;
}

// Triangle ctor normal (Trangle.fs:108-112): (e1 x e2) / |e1 x e2|
__device__ __forceinline__ D3 tri_normal_x(D3 e1, D3 e2)
{
    const D3 a = cross(e1, e2);
    const double al = sqrt(len2(a));
    return a / al;
}

// GetRandomInUnitSphere of the sphere sample (RayTracing.fs:261-266): the whole ball, no hemisphere test
static __device__ D3 random_in_unit_ball_x(const RngX &g, uint32_t dim)
{
    D3 p = mk3<double>(20., 20., 20.);
    uint32_t it = 0;
    while (dot(p, p) >= 1.0) {
        if (it >= MFX_REJECTION_CAP) return mk3<double>(0., 0., 0.);
        double u[4];
        rng_draw_x(g, dim, it++, u);
        p = mk3<double>(u[0], u[1], u[2]) * 2.0 - mk3<double>(1., 1., 1.);
    }
    return p;
}

// RandomInUnitDisk (RayTracing.fs:327-333): the loop runs at least once; draws (dim 0, iter 1, 2, ..)
static __device__ D3 random_in_unit_disk_x(const RngX &g)
{
    D3 p = mk3<double>(0., 0., 0.);
    double dt = 1.0;
    uint32_t it = 0;
    while (dt >= 1.0) {
        if (it >= MFX_REJECTION_CAP) return mk3<double>(0., 0., 0.);
        double u[4];
        rng_draw_x(g, MFX_DIM_CAMERA, 1u + it++, u);
        p = mk3<double>(u[0], u[1], 0.) * 2.0 - mk3<double>(1., 1., 0.);
        dt = dot(p, p);
    }
    return p;
}

// RayTraceCamera.GetRay(s, t) (RayTracing.fs:360-364); the Ray constructor normalises (:14-16)
__device__ __forceinline__ void lens_ray_x(const CamX &c, const LensX &lens, double s, double t, const RngX *g, D3 &origin, D3 &dir)
{
    D3 offset = mk3<double>(0., 0., 0.);
    if (g) {
        const D3 rd = random_in_unit_disk_x(*g) * lens.radius;
        offset = ld3(lens.u) * rd.x + ld3(lens.v) * rd.y;
    }
    const D3 target = (((ld3(c.topleft) + ld3(c.right) * s) + ld3(c.down) * t) - ld3(c.pos)) - offset;
    origin = ld3(c.pos) + offset;
    dir = normalize_x(target);
}

__device__ __forceinline__ D3 camera_ray_dir_x(const CamX &c, double u, double v)   // Camera.fs:134-139
{
    const D3 target = (ld3(c.topleft) + ld3(c.right) * u) + ld3(c.down) * v;
    return normalize_x(target - ld3(c.pos));
}
