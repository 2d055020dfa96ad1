/*
 * mafrix_oracle_sky.c -- CPU ORACLE for the sphere sample's integrator (test infrastructure, NOT the product;
 * PARITY UNPINNED, see mafrix_oracle.h).  Plain-C f64 restatement of GetColor and everything under it in
 * /root/reference/RenderTest/Sample/RayTracing.fs -- every function cites the lines it follows.  The vector
 * operators are those of EngineCore/Core/Point.fs (one rounding per written operation, -ffp-contract=off).
 *
 * What is ours, because the reference's is unreproducible by design: the random stream (Philox4x32-10 keyed on
 * pixel/sample/dimension, DESIGN.md "RNG") instead of `new System.Random()` / Random.Shared, and therefore the
 * Perlin tables (RayTracing.fs:81-85, filled from Random.Shared), which the caller supplies.
 * Deliberate deviations, stated once: (1) Schlick's Math.Pow(1-cosine, 5) (RayTracing.fs:280) is the product
 * x2 = x*x, x4 = x2*x2, x4*x -- it only feeds the comparison against a uniform of our own stream; (2) the wasted
 * `targ = n + GetRandomInUnitSphere()` of GetColor (:369, result unused, its own System.Random) draws nothing;
 * (3) the rejection loops are capped at 128 trips like the rest of the oracle; (4) MovingSphere (:210-253, not
 * used by RandomScene) and the ray's time are not restated.
 */
#include "mafrix_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

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
static inline V3 v_normalize(V3 a)                                                      /* Point.fs:52-56 */
{
    double l = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (l == 0.0) return v3(0, 0, 0);
    return v3(a.x / l, a.y / l, a.z / l);
}
static inline V3 ld(const double *p) { return v3(p[0], p[1], p[2]); }
static inline void st(double *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

#define SKY_CAP 128
#define SKY_TMIN 0.00001        /* RayTracing.fs:368 */
#define SKY_TMAX 10000000.

struct OrcSkyScene {
    int n, n_mats, width, height, max_depth;
    OrcPrim *spheres;
    OrcMaterial *mats;
    OrcLensCamera cam;
    double ranfloat[256];
    int32_t perm[768];
};

/* RayTraceCamera constructor, RayTracing.fs:346-358 */
void orc_camera_lens(const double lookfrom[3], const double lookat[3], const double vup[3], double vfov,
                     double aspect, double aperture, double focus_dist, OrcLensCamera *out)
{
    const double PI = 3.14159265358979323846;
    out->lens_radius = aperture / 2.0;
    double theta = vfov * PI / 180.;
    double half_height = tan(theta / 2.);
    double half_width = aspect * half_height;
    V3 origin = ld(lookfrom);
    V3 w = v_normalize(v_sub(ld(lookfrom), ld(lookat)));
    V3 u = v_normalize(v_cross(ld(vup), w));
    V3 v = v_normalize(v_cross(w, u));
    /* origin - focus_dist*half_width*u - focus_dist*half_height*v - focus_dist*w, left to right */
    V3 p = v_sub(v_sub(v_sub(origin, v_mul(u, focus_dist * half_width)), v_mul(v, focus_dist * half_height)), v_mul(w, focus_dist));
    st(out->origin, origin); st(out->lower_left, p);
    st(out->horizontal, v_mul(u, 2. * focus_dist * half_width));        /* ((2.*focus_dist)*half_width)*u */
    st(out->vertical, v_mul(v, 2. * focus_dist * half_height));
    st(out->u, u); st(out->v, v);
}

OrcSkyScene *orc_sky_create(const OrcPrim *spheres, int n, const OrcMaterial *mats, int n_mats,
                            const OrcLensCamera *cam, const double *ranfloat, const int32_t *perm,
                            int width, int height, int max_depth)
{
    OrcSkyScene *s = (OrcSkyScene *)calloc(1, sizeof(OrcSkyScene));
    s->n = n; s->n_mats = n_mats; s->width = width; s->height = height; s->max_depth = max_depth;
    s->spheres = (OrcPrim *)malloc(sizeof(OrcPrim) * (size_t)n);
    memcpy(s->spheres, spheres, sizeof(OrcPrim) * (size_t)n);
    s->mats = (OrcMaterial *)malloc(sizeof(OrcMaterial) * (size_t)n_mats);
    memcpy(s->mats, mats, sizeof(OrcMaterial) * (size_t)n_mats);
    s->cam = *cam;
    if (ranfloat) memcpy(s->ranfloat, ranfloat, sizeof(s->ranfloat));
    if (perm) memcpy(s->perm, perm, sizeof(s->perm));
    return s;
}
void orc_sky_destroy(OrcSkyScene *s) { if (s) { free(s->spheres); free(s->mats); free(s); } }

/* Ray(origin, direc): the constructor normalises (RayTracing.fs:14-21) */
typedef struct { V3 o, d; } Ray;
static inline Ray ray_make(V3 o, V3 direc) { Ray r; r.o = o; r.d = v_normalize(direc); return r; }

typedef struct { int hit; double t; V3 p, normal; int material, prim; } Rec;

/* Sphere.Hit, RayTracing.fs:188-207 */
static inline int sphere_hit(const OrcPrim *sp, const Ray *r, double tMin, double tMax, Rec *rec)
{
    const V3 center = ld(sp->v);
    const double radius = sp->v[3];
    V3 oc = v_sub(r->o, center);
    double a = v_dot(r->d, r->d);
    double b = 2.0 * v_dot(oc, r->d);
    double c = v_dot(oc, oc) - radius * radius;
    double disc = b * b - 4.0 * a * c;
    if (disc > 0) {
        double tmp = (-b - sqrt(disc)) / (2.0 * a);
        if (!(tmp < tMax && tmp > tMin)) {
            tmp = (-b + sqrt(disc)) / (2.0 * a);
            if (!(tmp < tMax && tmp > tMin)) return 0;
        }
        rec->hit = 1; rec->t = tmp;
        rec->p = v_add(r->o, v_mul(r->d, tmp));                     /* PointAtParameter, :21 */
        rec->normal = v_div(v_sub(rec->p, center), radius);
        rec->material = sp->material;
        return 1;
    }
    return 0;
}

/* ListHit, RayTracing.fs:256-258: every item is tested, Array.minBy keeps the FIRST minimal key
 * (hit ? t : tmax) */
static Rec list_hit(const OrcSkyScene *s, const Ray *r, double tmin, double tmax)
{
    Rec best; memset(&best, 0, sizeof(best)); best.prim = -1;
    double bestKey = 0.; int have = 0;
    for (int i = 0; i < s->n; i++) {
        Rec rec; memset(&rec, 0, sizeof(rec));
        int h = sphere_hit(&s->spheres[i], r, tmin, tmax, &rec);
        double key = h ? rec.t : tmax;
        if (!have || key < bestKey) { have = 1; bestKey = key; best = rec; best.prim = h ? i : -1; }
    }
    return best;
}

typedef struct { uint32_t pixel, sample, k0, k1; } Rng;
static void draw(const Rng *g, uint32_t dim, uint32_t iter, double u[4])
{
    uint32_t c[4] = { g->pixel, g->sample, dim, iter }, k[2] = { g->k0, g->k1 }, o[4];
    orc_philox4x32_10(c, k, o);
    for (int i = 0; i < 4; i++) u[i] = (double)o[i] * (1.0 / 4294967296.0);
}

/* GetRandomInUnitSphere, RayTracing.fs:261-266 (whole ball, no hemisphere test) */
static V3 random_in_unit_sphere(const Rng *g, uint32_t dim)
{
    V3 p = v3(20, 20, 20);
    uint32_t it = 0;
    while (v_dot(p, p) >= 1.0) {
        if (it >= SKY_CAP) return v3(0, 0, 0);
        double u[4];
        draw(g, dim, it++, u);
        p = v_sub(v_mul(v3(u[0], u[1], u[2]), 2.0), v3(1, 1, 1));
    }
    return p;
}

/* RandomInUnitDisk, RayTracing.fs:327-333: the loop always runs at least once */
static V3 random_in_unit_disk(const Rng *g)
{
    V3 p = v3(0, 0, 0);
    double dot = 1.0;
    uint32_t it = 0;
    while (dot >= 1.0) {
        if (it >= SKY_CAP) return v3(0, 0, 0);
        double u[4];
        draw(g, 0, 1 + it++, u);
        p = v_sub(v_mul(v3(u[0], u[1], 0), 2.0), v3(1, 1, 0));
        dot = v_dot(p, p);
    }
    return p;
}

static inline V3 reflect(V3 v, V3 n) { return v_sub(v, v_mul(n, 2.0 * v_dot(v, n))); }   /* :268 */
static int refract(V3 v, V3 n, double ni_over_nt, V3 *out)                               /* :269-276 */
{
    V3 uv = v_normalize(v);
    double dt = v_dot(uv, n);
    double disc = 1.0 - ni_over_nt * ni_over_nt * (1.0 - dt * dt);
    if (disc > 0) { *out = v_sub(v_mul(v_sub(v, v_mul(n, dt)), ni_over_nt), v_mul(n, sqrt(disc))); return 1; }
    *out = v3(0, 0, 0);
    return 0;
}
static double schlick(double cosine, double ref_idx)                                     /* :277-280 */
{
    double r0 = (1. - ref_idx) / (1. + ref_idx);
    double r1 = r0 * r0;
    double x = 1. - cosine, x2 = x * x, x4 = x2 * x2;
    return r1 + (1. - r1) * (x4 * x);
}

/* Texture.Value(0, 0, p): ConstantTexture :50-52, CheckerTexture :54-61, NoiseTexture + Perlin.Noise :81-99 */
static V3 texture_value(const OrcSkyScene *s, const OrcMaterial *m, V3 p)
{
    if (m->kind == ORC_LAMBERT_CHECKER) {
        double sines = sin(10. * p.x) * sin(10. * p.y) * sin(10. * p.z);
        if (sines < 0.) return v3(m->fuzz, m->ei, m->et);
        return ld(m->albedo);
    }
    if (m->kind == ORC_LAMBERT_NOISE) {
        int i = (int)(4. * p.x) & 255, j = (int)(4. * p.y) & 255, k = (int)(4. * p.z) & 255;
        double nz = s->ranfloat[s->perm[i] ^ s->perm[256 + j] ^ s->perm[512 + k]];
        return v_mul(v3(1, 1, 1), nz);
    }
    return ld(m->albedo);
}

/* Material.Scatter: Lambertian :282-290, Metal :291-299, Dielectric :300-325 */
static int scatter(const OrcSkyScene *s, const Ray *ray, const Rec *hit, const Rng *g, int k, V3 *att, Ray *scattered)
{
    const OrcMaterial *m = &s->mats[hit->material];
    if (m->kind == ORC_METAL) {
        double fuzz = m->fuzz < 1.0 ? m->fuzz : 1.0;
        V3 reflected = reflect(v_normalize(ray->d), hit->normal);
        *scattered = ray_make(hit->p, v_add(reflected, v_mul(random_in_unit_sphere(g, 1u + 2u * (uint32_t)k), fuzz)));
        *att = ld(m->albedo);
        return v_dot(scattered->d, hit->normal) > 0;
    }
    if (m->kind == ORC_DIELECTRIC) {
        const double ref_idx = m->ei;
        V3 reflected = reflect(ray->d, hit->normal);
        V3 outward; double ni_over_nt, cosine;
        if (v_dot(ray->d, hit->normal) > 0) {
            outward = v_neg(hit->normal); ni_over_nt = ref_idx; cosine = ref_idx * v_dot(ray->d, hit->normal);
        } else {
            outward = hit->normal; ni_over_nt = 1.0 / ref_idx; cosine = -v_dot(ray->d, hit->normal);
        }
        V3 ref_dir;
        int ok = refract(ray->d, outward, ni_over_nt, &ref_dir);
        double reflect_prob = ok ? schlick(cosine, ref_idx) : 1.0;
        double u[4];
        draw(g, 2u + 2u * (uint32_t)k, 0, u);
        *att = v3(1, 1, 1);
        *scattered = (u[0] < reflect_prob) ? ray_make(hit->p, reflected) : ray_make(hit->p, ref_dir);
        return 1;
    }
    /* Lambertian over a texture */
    V3 target = v_add(v_normalize(hit->normal), random_in_unit_sphere(g, 1u + 2u * (uint32_t)k));
    *scattered = ray_make(hit->p, target);
    *att = texture_value(s, m, hit->p);
    return 1;
}

/* GetColor, RayTracing.fs:367-382; `depth < 50` is max_depth */
static void get_color(const OrcSkyScene *s, const Ray *ray, int depth, const Rng *g, double rgb[3], uint64_t *rays)
{
    if (rays) (*rays)++;
    Rec hit = list_hit(s, ray, SKY_TMIN, SKY_TMAX);
    if (hit.hit) {
        V3 att; Ray scattered;
        int ok = scatter(s, ray, &hit, g, depth, &att, &scattered);
        if (depth < s->max_depth && ok) {
            double c[3];
            get_color(s, &scattered, depth + 1, g, c, rays);
            rgb[0] = c[0] * att.x; rgb[1] = c[1] * att.y; rgb[2] = c[2] * att.z;
        } else { rgb[0] = rgb[1] = rgb[2] = 0.; }
    } else {
        V3 unit = v_normalize(ray->d);
        double t = 0.5 * (unit.y + 1.0);
        V3 vec = v_add(v_mul(v3(1, 1, 1), 1.0 - t), v_mul(v3(0.5, 0.7, 1.0), t));
        rgb[0] = vec.x; rgb[1] = vec.y; rgb[2] = vec.z;
    }
}

/* RayTraceCamera.GetRay(s, t), RayTracing.fs:360-364; g == NULL: no lens sample (offset 0) */
static Ray camera_ray(const OrcLensCamera *c, double s, double t, const Rng *g)
{
    V3 offset = v3(0, 0, 0);
    if (g) {
        V3 rd = v_mul(random_in_unit_disk(g), c->lens_radius);
        offset = v_add(v_mul(ld(c->u), rd.x), v_mul(ld(c->v), rd.y));
    }
    V3 origin = ld(c->origin);
    V3 dir = v_sub(v_sub(v_add(v_add(ld(c->lower_left), v_mul(ld(c->horizontal), s)), v_mul(ld(c->vertical), t)), origin), offset);
    return ray_make(v_add(origin, offset), dir);
}

void orc_sky_list_hit(const OrcSkyScene *s, int n, const double *origins, const double *dirs, double tmin, double tmax,
                      int32_t *prim, double *t)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
        Ray r = ray_make(ld(origins + 3 * i), ld(dirs + 3 * i));
        Rec h = list_hit(s, &r, tmin, tmax);
        prim[i] = h.hit ? h.prim : -1;
        t[i] = h.hit ? h.t : 0.;
    }
}

void orc_sky_trace_primary(const OrcSkyScene *s, int n, const double *uv, int32_t *prim, double *t)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int r = 0; r < n; r++) {
        double u, v;
        if (uv) { u = uv[2 * r]; v = uv[2 * r + 1]; }
        else {
            int j = r / s->width, i = r - j * s->width;
            u = ((double)i + 0.5) / (double)s->width;
            v = ((double)j + 0.5) / (double)s->height;
        }
        Ray ray = camera_ray(&s->cam, u, v, NULL);
        Rec h = list_hit(s, &ray, SKY_TMIN, SKY_TMAX);
        prim[r] = h.hit ? h.prim : -1;
        t[r] = h.hit ? h.t : 0.;
    }
}

void orc_sky_trace_path(const OrcSkyScene *s, int px, int py, int sample, uint64_t seed, double *rgb_out, uint64_t *rays)
{
    Rng g = { (uint32_t)(py * s->width + px), (uint32_t)sample, (uint32_t)seed, (uint32_t)(seed >> 32) };
    double u4[4];
    draw(&g, 0, 0, u4);
    double u = ((double)px + u4[0]) / (double)s->width;            /* RayTracing.fs:450-451 (the commented driver loop) */
    double v = ((double)py + u4[1]) / (double)s->height;
    Ray ray = camera_ray(&s->cam, u, v, &g);
    get_color(s, &ray, 0, &g, rgb_out, rays);
}

/* The pixel loop of DoRayTrace (RayTracing.fs:444-455): col = sum over ns samples / ns; the sqrt / 255.99 /
 * vertical flip that follow (:456-460) belong to the display, not to the texture. */
void orc_sky_sample(const OrcSkyScene *s, int n, uint64_t seed, int first_sample, int threads, double *texture,
                    uint64_t *rays_out)
{
    uint64_t rays = 0;
    const int w = s->width, h = s->height;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : rays)
    for (int idx = 0; idx < w * h; idx++) {
        const int i = idx / h, j = idx - i * h;
        double cr = 0., cg = 0., cb = 0.;
        for (int sidx = 0; sidx < n; sidx++) {
            double c[3];
            uint64_t r = 0;
            orc_sky_trace_path(s, i, j, first_sample + sidx, seed, c, &r);
            rays += r;
            cr = cr + c[0]; cg = cg + c[1]; cb = cb + c[2];
        }
        double *o = texture + ((size_t)i * h + j) * 4;
        o[0] = cr / (double)n; o[1] = cg / (double)n; o[2] = cb / (double)n; o[3] = 1.0;
    }
    if (rays_out) *rays_out = rays;
}

/* What the pixel loop of DoRayTrace does with the mean colour before it reaches the window (RayTracing.fs:456-460):
 * col = (sqrt r, sqrt g, sqrt b); ir = int(255.99*col.r) ..; screen[i, (ny-1)-j] <- (ir, ig, ib).  texture is
 * Color[w,h] ([i,j] at (i*h+j)*4); rgba8 is row-major like Scene.PostProcessAndToScreenBuffer: x*4 + y*w*4, a = 255. */
void orc_sky_display_rgba8(const double *texture, int width, int height, uint8_t *rgba8)
{
    for (int i = 0; i < width; i++)
        for (int j = 0; j < height; j++) {
            const double *c = texture + ((size_t)i * height + j) * 4;
            uint8_t *o = rgba8 + ((size_t)(height - 1 - j) * width + i) * 4;
            for (int k = 0; k < 3; k++) {
                double v = 255.99 * sqrt(c[k]);
                o[k] = (v == v) ? (uint8_t)(int)v : 0;
            }
            o[3] = 255;
        }
}
