/*
 * mafrix_oracle.c -- CPU ORACLE (test infrastructure, NOT the product; PARITY UNPINNED,
 * see mafrix_oracle.h).  A plain-C, f64 restatement of the MafrixRender CPU path tracer.
 * Every function cites the reference file:line (paths under EngineCore/) it follows.
 *
 * Build:  gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared   (oracle/Makefile)
 */
#include "mafrix_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Core/Point.fs */
typedef struct { double x, y, z; } V3;

static inline V3 v3(double x, double y, double z) { V3 r = { x, y, z }; return r; }
static inline V3 v_sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }      /* Point.fs:31-32,64 */
static inline V3 v_add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }      /* Point.fs:60-61 */
static inline V3 v_neg(V3 a) { return v3(-a.x, -a.y, -a.z); }                           /* Point.fs:65 */
static inline V3 v_mul(V3 v, double a) { return v3(v.x * a, v.y * a, v.z * a); }        /* Point.fs:66-67 */
static inline V3 v_div(V3 v, double a) { return v3(v.x / a, v.y / a, v.z / a); }        /* Point.fs:68 */
static inline double v_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }    /* Point.fs:58 */
static inline V3 v_cross(V3 a, V3 v)                                                     /* Point.fs:57 */
{ return v3(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x); }
static inline double v_len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }         /* Point.fs:51 */
static inline double v_len(V3 a) { return sqrt(v_len2(a)); }                            /* Point.fs:50 */
static inline V3 v_normalize(V3 a)                                                      /* Point.fs:52-56 */
{
    double l = v_len(a);
    if (l == 0.0) return v3(0, 0, 0);
    return v3(a.x / l, a.y / l, a.z / l);
}
static inline double v_get(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* F# `min`/`max` on floats compile to System.Math.Min/Max (FSharp.Core prim-types, static
 * optimisation for float): NaN propagates and -0.0 orders below +0.0 (IEEE 754-2019 minimum). */
static inline double fs_min(double a, double b)
{
    if (a < b) return a;
    if (b < a) return b;
    if (a != a) return a;
    return signbit(a) ? a : b;
}
static inline double fs_max(double a, double b)
{
    if (a > b) return a;
    if (b > a) return b;
    if (a != a) return a;
    return signbit(a) ? b : a;
}

/* ------------------------------------------------------------------ Core/Color.fs (rgb only) */
typedef struct { double r, g, b; } Col;
static inline Col c3(double r, double g, double b) { Col c = { r, g, b }; return c; }
static inline Col c_mul(Col a, Col b) { return c3(a.r * b.r, a.g * b.g, a.b * b.b); }   /* Color.fs:13 */
static inline Col c_scale(double l, Col r) { return c3(l * r.r, l * r.g, l * r.b); }    /* Color.fs:14-15 */
static inline Col c_addf(Col l, double r) { return c3(l.r + r, l.g + r, l.b + r); }     /* Color.fs:16 */
static inline Col c_add(Col a, Col b) { return c3(a.r + b.r, a.g + b.g, a.b + b.b); }   /* Color.fs:17 */
static inline Col c_sub(Col a, Col b) { return c3(a.r - b.r, a.g - b.g, a.b - b.b); }   /* Color.fs:18 */
static inline Col c_divf(Col l, double r) { return c3(l.r / r, l.g / r, l.b / r); }     /* Color.fs:19 */
static inline Col c_div(Col a, Col b) { return c3(a.r / b.r, a.g / b.g, a.b / b.b); }   /* Color.fs:20 */

/* ------------------------------------------------------------------ Core/Aggregate.fs */
typedef struct { V3 pmin, pmax; } Bound;
static inline Bound bound2(V3 p1, V3 p2)                                                /* Aggregate.fs:9-11 */
{
    Bound b;
    b.pmin = v3(fs_min(p1.x, p2.x), fs_min(p1.y, p2.y), fs_min(p1.z, p2.z));
    b.pmax = v3(fs_max(p1.x, p2.x), fs_max(p1.y, p2.y), fs_max(p1.z, p2.z));
    return b;
}
static inline Bound bound_union_p(Bound b1, V3 p)                                       /* Aggregate.fs:58-61 */
{
    V3 p1 = v3(fs_min(b1.pmin.x, p.x), fs_min(b1.pmin.y, p.y), fs_min(b1.pmin.z, p.z));
    V3 p2 = v3(fs_max(b1.pmax.x, p.x), fs_max(b1.pmax.y, p.y), fs_max(b1.pmax.z, p.z));
    return bound2(p1, p2);
}
static inline Bound bound_union(Bound b1, Bound b2)                                     /* Aggregate.fs:62-65 */
{
    V3 p1 = v3(fs_min(b1.pmin.x, b2.pmin.x), fs_min(b1.pmin.y, b2.pmin.y), fs_min(b1.pmin.z, b2.pmin.z));
    V3 p2 = v3(fs_max(b1.pmax.x, b2.pmax.x), fs_max(b1.pmax.y, b2.pmax.y), fs_max(b1.pmax.z, b2.pmax.z));
    return bound2(p1, p2);
}
static inline int bound_max_extent(Bound b)                                             /* Aggregate.fs:29-36 */
{
    V3 d = v_sub(b.pmax, b.pmin);
    if (d.x > d.y && d.x > d.z) return 0;
    else if (d.y > d.z) return 1;
    else return 2;
}

/* ------------------------------------------------------------------ shapes */
typedef struct { V3 v0, v1, v2; double area; V3 normal; Bound bound; } Tri;            /* Trangle.fs:98-119 */

static Tri tri_make(V3 v0, V3 v1, V3 v2)                                                /* Trangle.fs:107-119 */
{
    Tri t;
    V3 e1 = v_sub(v1, v0), e2 = v_sub(v2, v0);
    V3 a = v_cross(e1, e2);
    double al = v_len(a);
    t.v0 = v0; t.v1 = v1; t.v2 = v2;
    t.area = al * 0.5;
    t.normal = v_div(a, al);
    t.bound = bound_union_p(bound2(v0, v1), v2);
    return t;
}

typedef struct {
    int    kind, material;
    Tri    t1, t2;          /* triangle: t1;  rect: t1,t2 (Rect.fs:17-25) */
    V3     center; double radius;   /* sphere (Sphere.fs:15-20)            */
    Bound  bound;
    double area;
} Prim;

typedef struct {
    int    hit;
    double t;
    V3     point, normal;
    int    material;
    int    prim, sub;
} HitRec;                                                                               /* HitRecord.fs:5-15 */

static const HitRec HIT_EMPTY = { 0, 0.0, { 0, 0, 0 }, { 0, 0, 0 }, 0, -1, 0 };

typedef struct { V3 o, d; } Ray;
static inline V3 ray_at(Ray r, double t) { return v_add(r.o, v_mul(r.d, t)); }          /* Ray.fs:8-10 */

/* Triangle.PreCalcu + Hit (Trangle.fs:120-155).  tMax is ignored (quirk Q2). */
static HitRec tri_hit(const Tri *tr, int material, Ray ray, double tMin)
{
    V3 e1 = v_sub(tr->v1, tr->v0);
    V3 e2 = v_sub(tr->v2, tr->v0);
    V3 s1 = v_cross(ray.d, e2);
    double divisor = v_dot(s1, e1);
    if (fabs(divisor) < 1e-6) return HIT_EMPTY;
    double inv = 1. / divisor;
    V3 d = v_sub(ray.o, tr->v0);
    double b1 = v_dot(d, s1) * inv;
    if (b1 < 0. || b1 > 1.) return HIT_EMPTY;
    V3 s2 = v_cross(d, e1);
    double b2 = v_dot(ray.d, s2) * inv;
    if (b2 < 0. || (b1 + b2) >= 1.) return HIT_EMPTY;
    double t = v_dot(e2, s2) * inv;
    if (t > tMin) {
        HitRec h;
        h.hit = 1; h.t = t; h.point = ray_at(ray, t); h.normal = tr->normal;
        h.material = material; h.prim = -1; h.sub = 0;
        return h;
    }
    return HIT_EMPTY;
}

/* Sphere.Hit (Sphere.fs:21-43) */
static HitRec sphere_hit(const Prim *sp, Ray r, double tMin, double tMax)
{
    V3 oc = v_sub(r.o, sp->center);
    double a = 1.;
    double b = 2.0 * v_dot(oc, r.d);
    double c = v_dot(oc, oc) - sp->radius * sp->radius;
    double disc = b * b - 4.0 * a * c;
    if (disc > 0) {
        double root = sqrt(disc);
        double q = (b < 0.) ? -0.5 * (b - root) : -0.5 * (b + root);
        double t0 = q;
        double t1 = c / q;
        double tmin = fs_min(t0, t1), tmax = fs_max(t0, t1);
        HitRec h;
        h.hit = 1; h.material = sp->material; h.prim = -1; h.sub = 0;
        if (tmin >= tMin && tmin < tMax) {
            V3 p = ray_at(r, tmin);
            h.t = tmin; h.point = p; h.normal = v_normalize(v_sub(p, sp->center));
            return h;
        } else if (tmax > tMin && tmax < tMax) {
            V3 p = ray_at(r, tmax);
            h.t = tmax; h.point = p; h.normal = v_normalize(v_sub(p, sp->center));
            return h;
        }
        return HIT_EMPTY;
    }
    return HIT_EMPTY;
}

/* IHitable.Hit dispatch; Rect.Hit (Rect.fs:26-31): tri1 if it hits, ELSE tri2 (quirk Q3). */
static HitRec prim_hit(const Prim *p, Ray r, double tMin, double tMax)
{
    if (p->kind == ORC_TRIANGLE) return tri_hit(&p->t1, p->material, r, tMin);
    if (p->kind == ORC_RECT) {
        HitRec h1 = tri_hit(&p->t1, p->material, r, tMin);
        if (h1.hit) return h1;
        HitRec h2 = tri_hit(&p->t2, p->material, r, tMin);
        h2.sub = 1;
        return h2;
    }
    return sphere_hit(p, r, tMin, tMax);
}

/* AABB.hit (IHitable.fs:18-54).  Divides by dir; `>= 0.` is true for -0.0 (quirk Q6).
 * *entry receives the final tmin (used only by the ordered counting traversal). */
static int aabb_hit(V3 pmin, V3 pmax, Ray r, double tMin, double tMax, double *entry)
{
    V3 o = r.o, dir = r.d;
    double tmin, tmax, tymin, tymax, tzmin, tzmax;
    if (dir.x >= 0.) { tmin = (pmin.x - o.x) / dir.x; tmax = (pmax.x - o.x) / dir.x; }
    else             { tmin = (pmax.x - o.x) / dir.x; tmax = (pmin.x - o.x) / dir.x; }
    if (dir.y >= 0.) { tymin = (pmin.y - o.y) / dir.y; tymax = (pmax.y - o.y) / dir.y; }
    else             { tymin = (pmax.y - o.y) / dir.y; tymax = (pmin.y - o.y) / dir.y; }
    if (tmin > tymax || tymin > tmax) return 0;
    tmin = (tymin > tmin) ? tymin : tmin;
    tmax = (tymax < tmax) ? tymax : tmax;
    if (dir.z >= 0.) { tzmin = (pmin.z - o.z) / dir.z; tzmax = (pmax.z - o.z) / dir.z; }
    else             { tzmin = (pmax.z - o.z) / dir.z; tzmax = (pmin.z - o.z) / dir.z; }
    if (tmin > tzmax || tzmin > tmax) return 0;
    tmin = (tzmin > tmin) ? tzmin : tmin;
    tmax = (tzmax < tmax) ? tzmax : tmax;
    if (entry) *entry = tmin;
    return tmin < tMax && tmax > tMin;
}

/* ------------------------------------------------------------------ scene */
struct OrcScene {
    int      n_prims, n_mats, n_slots;
    Prim    *prims;
    OrcMaterial *mats;
    OrcNode *nodes;         /* heap-indexed, 2N-1 slots (BvhNode.fs:26) */
    int32_t *indices;
    /* NewAreaLight (Light.fs:32-41) */
    Tri      lt1, lt2; double light_area; V3 light_n; Col light_color;
    /* PinholeCamera (Camera.fs:113-139) */
    V3       cam_pos, cam_topleft, cam_right, cam_down;
    int      width, height, max_depth, mode;
};

#define LEAF_NODE_COUNT 3                                                               /* BvhNode.fs:39 */

static Bound prim_bound(const Prim *p) { return p->bound; }

/* Bvh.InitNode (BvhNode.fs:32-37) */
static OrcNode init_node(const OrcScene *s, int start, int count)
{
    Bound b = prim_bound(&s->prims[s->indices[start]]);
    for (int i = 1; i < count; i++) b = bound_union(b, prim_bound(&s->prims[s->indices[start + i]]));
    OrcNode n;
    n.pmin[0] = b.pmin.x; n.pmin[1] = b.pmin.y; n.pmin[2] = b.pmin.z;
    n.pmax[0] = b.pmax.x; n.pmax[1] = b.pmax.y; n.pmax[2] = b.pmax.z;
    n.first = start; n.count = count;
    return n;
}

typedef struct { double key; int32_t idx; } SortItem;

/* Stable merge sort by key: Array.sortInPlaceBy is unstable in .NET (quirk Q9); the
 * restatement fixes the tie order to "stable". */
static void merge_sort(SortItem *a, SortItem *tmp, int n)
{
    if (n < 2) return;
    int h = n / 2;
    merge_sort(a, tmp, h);
    merge_sort(a + h, tmp, n - h);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, (size_t)n * sizeof(SortItem));
}

/* Bvh.Subdivide (BvhNode.fs:42-61) */
static void subdivide(OrcScene *s, int i, SortItem *buf, SortItem *tmp)
{
    OrcNode node = s->nodes[i];
    if (node.count > LEAF_NODE_COUNT) {
        Bound nb;
        nb.pmin = v3(node.pmin[0], node.pmin[1], node.pmin[2]);
        nb.pmax = v3(node.pmax[0], node.pmax[1], node.pmax[2]);
        int axis = bound_max_extent(nb);
        for (int k = 0; k < node.count; k++) {
            int32_t id = s->indices[node.first + k];
            Bound b = prim_bound(&s->prims[id]);
            V3 dig = v_mul(v_sub(b.pmax, b.pmin), 0.5);
            V3 p = v_add(b.pmin, dig);
            buf[k].key = v_get(p, axis);
            buf[k].idx = id;
        }
        merge_sort(buf, tmp, node.count);
        for (int k = 0; k < node.count; k++) s->indices[node.first + k] = buf[k].idx;
        int leftcount = node.count / 2;
        int li = i * 2 + 1, ri = i * 2 + 2;
        s->nodes[li] = init_node(s, node.first, leftcount);
        s->nodes[ri] = init_node(s, node.first + leftcount, node.count - leftcount);
        subdivide(s, li, buf, tmp);
        subdivide(s, ri, buf, tmp);
    }
}

/* Bvh.Build (BvhNode.fs:24-30) */
static void bvh_build(OrcScene *s)
{
    int n = s->n_prims;
    for (int i = 0; i < n; i++) s->indices[i] = i;
    memset(s->nodes, 0, (size_t)s->n_slots * sizeof(OrcNode));
    s->nodes[0] = init_node(s, 0, n);
    SortItem *buf = (SortItem *)malloc((size_t)n * sizeof(SortItem));
    SortItem *tmp = (SortItem *)malloc((size_t)n * sizeof(SortItem));
    subdivide(s, 0, buf, tmp);
    free(buf); free(tmp);
}

typedef struct { uint64_t nodes, prims; } RefCount;

/* Bvh.CheckHit (BvhNode.fs:62-82): exhaustive, both children always, no t-shrink. */
static HitRec check_hit(const OrcScene *s, Ray ray, double tMin, double tMax, int idx, RefCount *rc)
{
    const OrcNode *node = &s->nodes[idx];
    if (rc) rc->nodes++;
    if (aabb_hit(v3(node->pmin[0], node->pmin[1], node->pmin[2]),
                 v3(node->pmax[0], node->pmax[1], node->pmax[2]), ray, tMin, tMax, NULL)) {
        if (node->count > LEAF_NODE_COUNT) {
            HitRec l = check_hit(s, ray, tMin, tMax, idx * 2 + 1, rc);
            HitRec r = check_hit(s, ray, tMin, tMax, idx * 2 + 2, rc);
            if (l.hit && r.hit) return (l.t < r.t) ? l : r;
            else if (l.hit) return l;
            else return r;
        } else {
            /* Array.map Hit |> Array.minBy (hit ? t : tMax): FIRST minimal key wins */
            HitRec best = HIT_EMPTY; double bestkey = 0; int have = 0;
            for (int k = 0; k < node->count; k++) {
                int id = s->indices[node->first + k];
                HitRec h = prim_hit(&s->prims[id], ray, tMin, tMax);
                if (rc) rc->prims++;
                if (h.hit) h.prim = id;
                double key = h.hit ? h.t : tMax;
                if (!have || key < bestkey) { best = h; bestkey = key; have = 1; }
            }
            return best;
        }
    }
    return HIT_EMPTY;
}

static HitRec bvh_hit(const OrcScene *s, Ray ray, double tMin, double tMax, RefCount *rc) /* BvhNode.fs:83 */
{
    return check_hit(s, ray, tMin, tMax, 0, rc);
}

/* ---- counting-only traversal of the SAME tree: ordered descent + t-shrink (closest) or
 * first-hit early-out (shadow).  Defines the algorithmic record counts of SURVEY 8(d).
 * Children are fetched as a pair (2 node records per interior visit, +1 for the root). */
typedef struct { uint64_t nodes, tris, spheres; } OrdCount;

static void count_leaf(const OrcScene *s, const OrcNode *node, Ray ray, double tMin, double tMax,
                       double *best, int *anyhit, OrdCount *oc)
{
    for (int k = 0; k < node->count; k++) {
        const Prim *p = &s->prims[s->indices[node->first + k]];
        if (p->kind == ORC_SPHERE) oc->spheres++;
        else if (p->kind == ORC_RECT) oc->tris += 2;
        else oc->tris++;
        HitRec h = prim_hit(p, ray, tMin, tMax);
        if (h.hit && h.t < tMax) { *anyhit = 1; if (h.t < *best) *best = h.t; }
    }
}

static void count_ordered(const OrcScene *s, Ray ray, double tMin, double tMax, int shadow, OrdCount *oc)
{
    int stack[64]; double stack_t[64]; int sp = 0;
    double best = tMax; int anyhit = 0; double e;
    const OrcNode *root = &s->nodes[0];
    oc->nodes++;
    if (!aabb_hit(v3(root->pmin[0], root->pmin[1], root->pmin[2]),
                  v3(root->pmax[0], root->pmax[1], root->pmax[2]), ray, tMin, tMax, &e)) return;
    int cur = 0;
    for (;;) {
        const OrcNode *node = &s->nodes[cur];
        if (node->count > LEAF_NODE_COUNT) {
            int c[2] = { cur * 2 + 1, cur * 2 + 2 }; int h[2]; double te[2] = { 0, 0 };
            oc->nodes += 2;
            for (int k = 0; k < 2; k++) {
                const OrcNode *ch = &s->nodes[c[k]];
                h[k] = aabb_hit(v3(ch->pmin[0], ch->pmin[1], ch->pmin[2]),
                                v3(ch->pmax[0], ch->pmax[1], ch->pmax[2]), ray, tMin, best, &te[k]);
            }
            if (h[0] && h[1]) {
                int nearc = (te[1] < te[0]) ? 1 : 0;
                stack[sp] = c[1 - nearc]; stack_t[sp] = te[1 - nearc]; sp++;
                cur = c[nearc];
                continue;
            } else if (h[0]) { cur = c[0]; continue; }
            else if (h[1]) { cur = c[1]; continue; }
        } else {
            count_leaf(s, node, ray, tMin, tMax, &best, &anyhit, oc);
            if (shadow && anyhit) return;
        }
        /* pop, culling entries whose entry distance is now behind the best hit */
        for (;;) {
            if (sp == 0) return;
            sp--;
            if (stack_t[sp] < best) { cur = stack[sp]; break; }
        }
    }
}

/* ------------------------------------------------------------------ RNG (DESIGN.md "RNG") */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        philox_round(c, k0, k1);
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* Stream layout: counter = (pixel = j*width+i, sample, dim, iteration), key = seed.
 * dim 0 = camera jitter; vertex k (0 = primary hit): dim 1+2k = BSDF/scatter rejection
 * loop (iteration = loop trip), dim 2+2k = light sample.  uniform = u32 * 2^-32. */
typedef struct { uint32_t pixel, sample, k0, k1; } Rng;
#define DIM_CAMERA 0u
#define DIM_BSDF(k)  (1u + 2u * (uint32_t)(k))
#define DIM_LIGHT(k) (2u + 2u * (uint32_t)(k))
#define REJECTION_CAP 128

static inline void rng_draw(const Rng *g, uint32_t dim, uint32_t iter, double u[4])
{
    uint32_t c[4] = { g->pixel, g->sample, dim, iter }, k[2] = { g->k0, g->k1 }, o[4];
    orc_philox4x32_10(c, k, o);
    for (int i = 0; i < 4; i++) u[i] = (double)o[i] * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------ Materials/Material.fs */
/* GetRandomInUnitSphere (Material.fs:9-14): rejection loop, unnormalised result.
 * The loop is capped at REJECTION_CAP trips (probability 0.738^128 ~ 1e-17); the cap
 * returns nm itself.  The kernels apply the identical cap. */
static V3 random_in_unit_sphere(V3 nm, const Rng *g, uint32_t dim)
{
    V3 p = v3(20, 20, 20);
    uint32_t it = 0;
    while (v_dot(p, p) >= 1.0 || v_dot(nm, p) <= 0.) {
        if (it >= REJECTION_CAP) return nm;
        double u[4];
        rng_draw(g, dim, it++, u);
        p = v_sub(v_mul(v3(u[0], u[1], u[2]), 2.0), v3(1, 1, 1));
    }
    return p;
}

static inline V3 reflect(V3 v, V3 n) { return v_sub(v, v_mul(n, 2.0 * v_dot(v, n))); }  /* Material.fs:16 */

static int refract(V3 v, V3 n, double ni_over_nt, V3 *out)                              /* Material.fs:17-24 */
{
    V3 uv = v_normalize(v);
    double dt = v_dot(uv, n);
    double disc = 1.0 - ni_over_nt * ni_over_nt * (1.0 - dt * dt);
    if (disc > 0) {
        *out = v_sub(v_mul(v_sub(v, v_mul(n, dt)), ni_over_nt), v_mul(n, sqrt(disc)));
        return 1;
    }
    *out = reflect(v, n);
    return 0;
}

#define ORC_PI 3.14159265358979323846
static const double INVPI = 1. / ORC_PI;                                                /* Material.fs:26 */
static const double TWOPI = 2. * ORC_PI;                                                /* Material.fs:27 */

static double fresnel_dielectric(double eta_i, double eta_t, double cosi)               /* Material.fs:74-96 */
{
    double ei, et;
    if (cosi > 0.) { ei = eta_i; et = eta_t; } else { ei = eta_t; et = eta_i; }
    double sint = ei / et * sqrt(fs_max(0., 1. - cosi * cosi));
    if (sint >= 1.) return 1.0;
    double cost = sqrt(fs_max(0., 1. - sint * sint));
    double ci = fabs(cosi);
    double rparl = ((et * ci) - (ei * cost)) / ((et * ci) + (ei * cost));
    double rperp = ((ei * ci) - (et * cost)) / ((ei * ci) + (et * cost));
    return (rparl * rparl + rperp * rperp) / 2.;
}

/* ------------------------------------------------------------------ Lights/Light.fs */
/* Triangle.SamplePoint (Trangle.fs:157-169) with tu,tv supplied */
static V3 tri_sample_point(const Tri *t, double tu, double tv)
{
    double u, v;
    if (tu + tv > 1.) { u = 1. - tu; v = 1. - tv; } else { u = tu; v = tv; }
    V3 e1 = v_sub(t->v1, t->v0), e2 = v_sub(t->v2, t->v0);
    double sq = sqrt(1. - u);
    double s1 = 1. - sq;
    double s2 = v * sq;
    return v_add(v_add(t->v0, v_mul(e1, s1)), v_mul(e2, s2));
}

/* NewAreaLight.GetDirection (Light.fs:42-47) via Rect.SamplePoint (Rect.fs:33-38) */
static void light_get_direction(const OrcScene *s, V3 p, const Rng *g, int k, double *dist, V3 *toLight)
{
    double u[4];
    rng_draw(g, DIM_LIGHT(k), 0, u);
    V3 lp = (u[0] < 0.5) ? tri_sample_point(&s->lt1, u[1], u[2]) : tri_sample_point(&s->lt2, u[1], u[2]);
    *toLight = v_sub(lp, p);
    *dist = v_len(*toLight);
}

/* NewAreaLight.L (Light.fs:48-56): unnormalised toLight, one-sided (quirk Q7) */
static Col light_L(const OrcScene *s, V3 toLight)
{
    double cos_o = v_dot(toLight, s->light_n);
    if (cos_o < 0.) {
        double dist = v_len2(toLight);
        double solidAngle = fabs(cos_o) * s->light_area / dist;
        return c_scale(solidAngle, s->light_color);
    }
    return c3(0, 0, 0);
}

/* ------------------------------------------------------------------ integrators */
typedef struct {
    OrcStats *st;          /* may be NULL */
    int count_ordered;
} Ctx;

static HitRec closest_query(const OrcScene *s, Ray ray, int depth, Ctx *cx)
{
    RefCount rc = { 0, 0 };
    HitRec h = bvh_hit(s, ray, 1e-6, 99999999., cx->st ? &rc : NULL);                   /* Integrators.fs:108 */
    if (cx->st) {
        cx->st->ref_nodes += rc.nodes; cx->st->ref_prims += rc.prims;
        if (depth >= 0) cx->st->closest_rays++; else cx->st->wasted_rays++;
        if (cx->count_ordered && depth >= 0) {
            OrdCount oc = { 0, 0, 0 };
            count_ordered(s, ray, 1e-6, 99999999., 0, &oc);
            cx->st->ord_rays[0]++; cx->st->ord_nodes[0] += oc.nodes;
            cx->st->ord_tris[0] += oc.tris; cx->st->ord_spheres[0] += oc.spheres;
        }
    }
    return h;
}

/* SingleDirectLightIntegrator.Eval / VisibilityTest (Integrators.fs:21-29,41-52) */
static Col direct_light(const OrcScene *s, const HitRec *hit, const Rng *g, int k, Ctx *cx)
{
    double dist; V3 toLight;
    light_get_direction(s, hit->point, g, k, &dist, &toLight);
    V3 unit = v_div(toLight, dist);
    Ray sr; sr.o = hit->point; sr.d = unit;
    RefCount rc = { 0, 0 };
    HitRec sh = bvh_hit(s, sr, 1e-6, dist - 1e-6, cx->st ? &rc : NULL);
    if (cx->st) {
        cx->st->ref_nodes += rc.nodes; cx->st->ref_prims += rc.prims; cx->st->shadow_rays++;
        if (cx->count_ordered) {
            OrdCount oc = { 0, 0, 0 };
            count_ordered(s, sr, 1e-6, dist - 1e-6, 1, &oc);
            cx->st->ord_rays[1]++; cx->st->ord_nodes[1] += oc.nodes;
            cx->st->ord_tris[1] += oc.tris; cx->st->ord_spheres[1] += oc.spheres;
        }
    }
    if (sh.hit) return c3(0, 0, 0);
    Col l = light_L(s, toLight);
    return c_scale(v_dot(unit, hit->normal), l);
}

/* PathIntegrator.TraceRay (Integrators.fs:107-138) -- mode A */
static Col trace_ray_a(const OrcScene *s, Ray ray, int depth, const Rng *g, Ctx *cx)
{
    HitRec hit = closest_query(s, ray, depth, cx);
    if (hit.hit && depth >= 0) {
        int k = s->max_depth - depth;
        const OrcMaterial *m = &s->mats[hit.material];
        /* material.GetBxdf(): Lambertian/Metal -> LambertianBrdf(a); SpecularTransmission ->
         * LambertianBrdf(Color()) (Material.fs:52,68,121) */
        Col a = (m->kind == ORC_SPECTRANS) ? c3(0, 0, 0) : c3(m->albedo[0], m->albedo[1], m->albedo[2]);
        /* LambertianBrdf.SampleF (Material.fs:33-36) */
        V3 wi = v_normalize(random_in_unit_sphere(hit.normal, g, DIM_BSDF(k)));
        double ei = v_dot(hit.normal, wi);
        Col col = c_scale(TWOPI, c_scale(ei, c_scale(INVPI, a)));
        double pdf = 1.;
        /* lightInteg.Eval: pdf_li = 1/area (Light.fs:57-59) */
        double pdf_li = 1. / s->light_area;
        Col l = direct_light(s, &hit, g, k, cx);
        Ray r; r.o = hit.point; r.d = wi;
        Col li = trace_ray_a(s, r, depth - 1, g, cx);
        return c_divf(c_mul(c_add(c_divf(l, pdf_li), li), col), pdf);                   /* Integrators.fs:136 */
    }
    return c3(0, 0, 0);
}

/* NewPathTracer.TraceRay (PathTracer.fs:23-43) -- mode B */
static Col trace_ray_b(const OrcScene *s, Ray ray, int depth, const Rng *g, Ctx *cx)
{
    HitRec hit = closest_query(s, ray, depth, cx);
    if (hit.hit && depth >= 0) {
        int k = s->max_depth - depth;
        const OrcMaterial *m = &s->mats[hit.material];
        Col albedo = c3(m->albedo[0], m->albedo[1], m->albedo[2]);
        Col col; Ray r; r.o = hit.point;
        if (m->kind == ORC_LAMBERT) {                                                   /* Material.fs:41-46 */
            r.d = v_normalize(random_in_unit_sphere(hit.normal, g, DIM_BSDF(k)));
            col = c_scale(INVPI, albedo);
        } else if (m->kind == ORC_METAL) {                                              /* Material.fs:61-65 */
            double fuzz = (m->fuzz < 1.0) ? m->fuzz : 1.0;
            V3 reflected = reflect(ray.d, hit.normal);
            V3 rnd = random_in_unit_sphere(hit.normal, g, DIM_BSDF(k));
            r.d = v_normalize(v_add(reflected, v_mul(rnd, fuzz)));
            col = albedo;
        } else {                                                                        /* Material.fs:103-118 */
            V3 dir = v_neg(ray.d);
            double cosi = v_dot(dir, hit.normal);
            double ei, et;
            if (cosi > 0.) { ei = m->ei; et = m->et; } else { ei = m->et; et = m->ei; }
            V3 refr;
            int isRefract = refract(dir, hit.normal, ei / et, &refr);
            if (isRefract) {
                double F = fresnel_dielectric(m->ei, m->et, cosi);
                Col one_minus_f = c_sub(c3(1, 1, 1), c3(F, F, F));
                col = c_divf(c_mul(c_scale((et * et) / (ei * ei), one_minus_f), albedo),
                             fabs(v_dot(refr, hit.normal)));
            } else col = c3(0, 0, 0);
            r.d = refr;
        }
        Col t = trace_ray_b(s, r, depth - 1, g, cx);
        Col shade;
        if (m->kind == ORC_LAMBERT) {                                                   /* Material.fs:47-50 */
            double ei = v_dot(hit.normal, r.d);
            shade = c_scale(ei, c_scale(ORC_PI, c_scale(2., t)));
        } else shade = t;                                                               /* Material.fs:66,119 */
        Col l = direct_light(s, &hit, g, k, cx);                                        /* PathTracer.fs:14-22 */
        return c_add(c_mul(l, col), c_mul(col, shade));                                 /* PathTracer.fs:40-41 */
    }
    return c3(0, 0, 0);
}

/* PinholeCamera.GetRay (Camera.fs:134-139) */
static Ray camera_get_ray(const OrcScene *s, double u, double v)
{
    V3 target = v_add(v_add(s->cam_topleft, v_mul(s->cam_right, u)), v_mul(s->cam_down, v));
    Ray r; r.o = s->cam_pos; r.d = v_normalize(v_sub(target, s->cam_pos));
    return r;
}

static Col trace_path(const OrcScene *s, int i, int j, int sample, uint64_t seed, Ctx *cx)
{
    Rng g; g.pixel = (uint32_t)(j * s->width + i); g.sample = (uint32_t)sample;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    double u4[4];
    rng_draw(&g, DIM_CAMERA, 0, u4);
    double u = ((double)i + u4[0]) / (double)s->width;                                  /* Integrators.fs:167 */
    double v = ((double)j + u4[1]) / (double)s->height;                                 /* Integrators.fs:168 */
    Ray ray = camera_get_ray(s, u, v);
    return (s->mode == ORC_MODE_A) ? trace_ray_a(s, ray, s->max_depth, &g, cx)
                                   : trace_ray_b(s, ray, s->max_depth, &g, cx);
}

/* ------------------------------------------------------------------ public API */
void orc_camera_pinhole(const double pos[3], const double dir[3], double fov, double aspect, double out[12])
{
    /* CameraCoordinate(dir) (Camera.fs:96-104) */
    V3 d = v_normalize(v3(dir[0], dir[1], dir[2]));
    V3 up0 = v_normalize(v3(0, 1, 0));
    V3 hori = v_cross(d, v_normalize(up0));
    V3 vert = v_cross(hori, d);
    /* PinholeCamera ctor (Camera.fs:122-133); effective FOV = fov/2 (quirk Q5) */
    double h = tan(0.5 * fov * ORC_PI / 360.);
    double vv = h / aspect;
    V3 up = v_mul(vert, vv), right = v_mul(hori, h), down = v_neg(up);
    V3 p = v3(pos[0], pos[1], pos[2]);
    /* TopLeft: pos + dist*forward - 0.5*right + 0.5*up (Camera.fs:110-111) */
    V3 tl = v_add(v_sub(v_add(p, v_mul(d, 0.5)), v_mul(right, 0.5)), v_mul(up, 0.5));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
    out[3] = tl.x; out[4] = tl.y; out[5] = tl.z;
    out[6] = right.x; out[7] = right.y; out[8] = right.z;
    out[9] = down.x; out[10] = down.y; out[11] = down.z;
}

static V3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }

OrcScene *orc_scene_create(const OrcPrim *prims, int n, const OrcMaterial *mats, int m,
                           const double lp[12], const double ln[3], const double lc[3],
                           const double cam[12], int width, int height, int max_depth, int mode)
{
    OrcScene *s = (OrcScene *)calloc(1, sizeof(OrcScene));
    s->n_prims = n; s->n_mats = m; s->n_slots = n * 2 - 1;
    s->prims = (Prim *)calloc((size_t)n, sizeof(Prim));
    s->mats = (OrcMaterial *)malloc((size_t)(m > 0 ? m : 1) * sizeof(OrcMaterial));
    memcpy(s->mats, mats, (size_t)m * sizeof(OrcMaterial));
    s->nodes = (OrcNode *)calloc((size_t)s->n_slots, sizeof(OrcNode));
    s->indices = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        Prim *p = &s->prims[i]; const double *v = prims[i].v;
        p->kind = prims[i].kind; p->material = prims[i].material;
        if (p->kind == ORC_TRIANGLE) {
            p->t1 = tri_make(ld3(v), ld3(v + 3), ld3(v + 6));
            p->bound = p->t1.bound; p->area = p->t1.area;
        } else if (p->kind == ORC_RECT) {                                               /* Rect.fs:17-25 */
            p->t1 = tri_make(ld3(v), ld3(v + 3), ld3(v + 6));
            p->t2 = tri_make(ld3(v), ld3(v + 6), ld3(v + 9));
            p->bound = bound_union(p->t1.bound, p->t2.bound);
            p->area = p->t1.area + p->t2.area;
        } else {                                                                        /* Sphere.fs:15-20 */
            p->center = ld3(v); p->radius = v[3];
            V3 rv = v3(v[3], v[3], v[3]);
            p->bound = bound2(v_sub(p->center, rv), v_add(p->center, rv));
        }
    }
    /* NewAreaLight (Light.fs:36-41): rect = Rect(p0,p1,p2,p3,0) */
    s->lt1 = tri_make(ld3(lp), ld3(lp + 3), ld3(lp + 6));
    s->lt2 = tri_make(ld3(lp), ld3(lp + 6), ld3(lp + 9));
    s->light_area = s->lt1.area + s->lt2.area;
    s->light_n = ld3(ln);
    s->light_color = c3(lc[0], lc[1], lc[2]);
    s->cam_pos = ld3(cam); s->cam_topleft = ld3(cam + 3); s->cam_right = ld3(cam + 6); s->cam_down = ld3(cam + 9);
    s->width = width; s->height = height; s->max_depth = max_depth; s->mode = mode;
    bvh_build(s);
    return s;
}

void orc_scene_destroy(OrcScene *s)
{
    if (!s) return;
    free(s->prims); free(s->mats); free(s->nodes); free(s->indices); free(s);
}

int orc_scene_node_slots(const OrcScene *s) { return s->n_slots; }

void orc_scene_get_bvh(const OrcScene *s, OrcNode *nodes, int32_t *indices)
{
    memcpy(nodes, s->nodes, (size_t)s->n_slots * sizeof(OrcNode));
    memcpy(indices, s->indices, (size_t)s->n_prims * sizeof(int32_t));
}

void orc_scene_set_bvh(OrcScene *s, const OrcNode *nodes, const int32_t *indices)
{
    memcpy(s->nodes, nodes, (size_t)s->n_slots * sizeof(OrcNode));
    memcpy(s->indices, indices, (size_t)s->n_prims * sizeof(int32_t));
}

void orc_bvh_hit(const OrcScene *s, int n, const double *o, const double *d, double tmin, double tmax,
                 int32_t *prim, int32_t *sub, double *t, double *point, double *normal)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        Ray r; r.o = ld3(o + 3 * i); r.d = ld3(d + 3 * i);
        HitRec h = bvh_hit(s, r, tmin, tmax, NULL);
        prim[i] = h.hit ? h.prim : -1;
        if (sub) sub[i] = h.hit ? h.sub : 0;
        t[i] = h.hit ? h.t : 0.0;
        if (point) { point[3 * i] = h.point.x; point[3 * i + 1] = h.point.y; point[3 * i + 2] = h.point.z; }
        if (normal) { normal[3 * i] = h.normal.x; normal[3 * i + 1] = h.normal.y; normal[3 * i + 2] = h.normal.z; }
    }
}

void orc_trace_primary(const OrcScene *s, int n, const double *uv, int32_t *prim, double *t)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int r = 0; r < n; r++) {
        double u, v;
        if (uv) { u = uv[2 * r]; v = uv[2 * r + 1]; }
        else {
            int i = r % s->width, j = r / s->width;
            u = ((double)i + 0.5) / (double)s->width;
            v = ((double)j + 0.5) / (double)s->height;
        }
        Ray ray = camera_get_ray(s, u, v);
        HitRec h = bvh_hit(s, ray, 1e-6, 99999999., NULL);
        prim[r] = h.hit ? h.prim : -1;
        t[r] = h.hit ? h.t : 0.0;
    }
}

static void stats_add(OrcStats *a, const OrcStats *b)
{
    a->closest_rays += b->closest_rays; a->wasted_rays += b->wasted_rays; a->shadow_rays += b->shadow_rays;
    a->ref_nodes += b->ref_nodes; a->ref_prims += b->ref_prims;
    for (int c = 0; c < 2; c++) {
        a->ord_rays[c] += b->ord_rays[c]; a->ord_nodes[c] += b->ord_nodes[c];
        a->ord_tris[c] += b->ord_tris[c]; a->ord_spheres[c] += b->ord_spheres[c];
    }
}

/* PixelIntegrator.Sample (Integrators.fs:160-172): x-major pixel order, Color[w,h]. */
void orc_sample(const OrcScene *s, int n, uint64_t seed, int first_sample,
                int x0, int y0, int x1, int y1, int threads,
                double *texture, OrcStats *stats, int count_ord)
{
    int bw = x1 - x0, bh = y1 - y0;
    long total = (long)bw * bh;
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    if (stats) memset(stats, 0, sizeof(*stats));
#pragma omp parallel num_threads(nt)
    {
        OrcStats local; memset(&local, 0, sizeof(local));
        Ctx cx; cx.st = stats ? &local : NULL; cx.count_ordered = count_ord;
#pragma omp for schedule(dynamic, 64)
        for (long p = 0; p < total; p++) {
            int i = x0 + (int)(p / bh), j = y0 + (int)(p % bh);
            Col color = c3(0, 0, 0);
            for (int sidx = 0; sidx < n; sidx++)
                color = c_add(color, trace_path(s, i, j, first_sample + sidx, seed, &cx));
            Col out = c_divf(color, (double)n);
            double *px = texture + ((size_t)i * s->height + j) * 4;
            px[0] = out.r; px[1] = out.g; px[2] = out.b; px[3] = 1.0;
        }
        if (stats) {
#pragma omp critical
            stats_add(stats, &local);
        }
    }
}

void orc_trace_path(const OrcScene *s, int px, int py, int sample, uint64_t seed, double *rgb)
{
    Ctx cx; cx.st = NULL; cx.count_ordered = 0;
    Col c = trace_path(s, px, py, sample, seed, &cx);
    rgb[0] = c.r; rgb[1] = c.g; rgb[2] = c.b;
}

void orc_film_add_sample(double *sum, const double *frame, double *target, int n_pixels, double frame_count)
{
    for (int p = 0; p < n_pixels; p++) {
        for (int c = 0; c < 3; c++) {
            double v = sum[4 * p + c] + frame[4 * p + c];                               /* Film.fs:21 */
            sum[4 * p + c] = v;
            target[4 * p + c] = v / frame_count;                                        /* Film.fs:23 */
        }
        /* alpha: min(a+a,1) then Color/float -> 1.0 (Color.fs:17,19) */
        double a = sum[4 * p + 3] + frame[4 * p + 3];
        sum[4 * p + 3] = a < 1.0 ? a : 1.0;
        target[4 * p + 3] = 1.0;
    }
}

static inline double clamp01(double x) { return x < 0. ? 0. : (x > 1. ? 1. : x); }      /* Scene.fs:273 */

void orc_tonemap_rgba8(const double *texture, int width, int height, uint8_t *rgba8)
{
    const double a = 2.51, b = 0.03, c = 2.43, d = 0.59, e = 0.14;                      /* Scene.fs:283-287 */
    for (int x = 0; x < width; x++)
        for (int y = 0; y < height; y++) {
            const double *px = texture + ((size_t)x * height + y) * 4;
            Col col = c3(px[0], px[1], px[2]);
            /* ((x*(a*x+b))/(x*(c*x+d)+e)) with Color operators (Scene.fs:288) */
            Col num = c_mul(col, c_addf(c_scale(a, col), b));
            Col den = c_addf(c_mul(col, c_addf(c_scale(c, col), d)), e);
            Col q = c_div(num, den);
            q = c3(clamp01(q.r), clamp01(q.g), clamp01(q.b));
            q = c3(sqrt(q.r), sqrt(q.g), sqrt(q.b));                                    /* Scene.fs:320 */
            int ir = (int)(255.99 * q.r), ig = (int)(255.99 * q.g), ib = (int)(255.99 * q.b);
            size_t o = (size_t)x * 4 + (size_t)y * (size_t)width * 4;                   /* Scene.fs:326-329 */
            rgba8[o] = (uint8_t)ir; rgba8[o + 1] = (uint8_t)ig; rgba8[o + 2] = (uint8_t)ib; rgba8[o + 3] = 255;
        }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
