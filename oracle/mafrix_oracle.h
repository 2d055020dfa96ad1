/*
 * mafrix_oracle.h -- CPU ORACLE for the MafrixRender path tracer.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference (NAIVEddd/MafrixRaytracing, F#/.NET 6) ships no tests,
 * golden vectors or fixtures for this path and cannot be compiled or run in this image
 * (no dotnet/mono/fsc).  This file is therefore a line-by-line *restatement* of the F#
 * sources (each function cites the file:line it follows), pinned only by analytic
 * known-answer tests minted in tests/ (see DESIGN.md "Oracle").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libmafrix_cuda) never links or calls it.
 *
 * All arithmetic is IEEE f64, compiled with -ffp-contract=off (RyuJIT never contracts
 * a*b+c), in the operation order of the F# expressions.
 */
#ifndef MAFRIX_ORACLE_H
#define MAFRIX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_TRIANGLE = 0, ORC_RECT = 1, ORC_SPHERE = 2 };
enum { ORC_LAMBERT = 0, ORC_METAL = 1, ORC_SPECTRANS = 2,
       /* sphere sample only (RenderTest/Sample/RayTracing.fs:300-325, :54-61, :96-99) */
       ORC_DIELECTRIC = 3, ORC_LAMBERT_CHECKER = 4, ORC_LAMBERT_NOISE = 5 };
enum { ORC_MODE_A = 0 /* PathIntegrator, Integrators.fs:96-141 */,
       ORC_MODE_B = 1 /* NewPathTracer,  PathTracer.fs:13-46   */ };

/* One IHitable.  Triangle: v[0..8] = v0,v1,v2.  Rect: v[0..11] = v0..v3.
 * Sphere: v[0..2] = center, v[3] = radius. */
typedef struct {
    int32_t kind;
    int32_t material;
    double  v[12];
} OrcPrim;

/* One IMaterial. lambert: albedo.  metal: albedo, fuzz.  spec-trans: albedo = T, ei, et. */
typedef struct {
    int32_t kind;
    int32_t pad;
    double  albedo[3];
    double  fuzz;
    double  ei;
    double  et;
} OrcMaterial;

/* BvhNode (BvhNode.fs:11-17): bound + first + count, heap-indexed. */
typedef struct {
    double  pmin[3];
    double  pmax[3];
    int32_t first;
    int32_t count;
} OrcNode;

/* Ray / record counters.  Class 0 = closest-hit queries, class 1 = shadow queries.
 * ref_*: work done by the reference's exhaustive CheckHit (BvhNode.fs:62-82).
 * ord_*: records touched by an ordered, t-shrinking (closest) / first-hit early-out
 *        (shadow) traversal of the same tree -- the "algorithmic bytes" of SURVEY 8(d):
 *        B_ray = 32*nodes + 48*tris + 16*spheres + 64.  A Rect counts as 2 triangles. */
typedef struct {
    uint64_t closest_rays;      /* bvh.Hit closest queries at depth >= 0 (useful)            */
    uint64_t wasted_rays;       /* the depth -1 closest query of Integrators.fs:108-109      */
    uint64_t shadow_rays;       /* shadow queries                                            */
    uint64_t ref_nodes;         /* AABB tests by the exhaustive traversal, all queries       */
    uint64_t ref_prims;         /* IHitable.Hit calls by the exhaustive traversal            */
    uint64_t ord_rays[2];
    uint64_t ord_nodes[2];
    uint64_t ord_tris[2];
    uint64_t ord_spheres[2];
} OrcStats;

typedef struct OrcScene OrcScene;

/* PinholeCamera ctor (Camera.fs:96-133): derives pos/topleft/right/down (12 doubles). */
void orc_camera_pinhole(const double pos[3], const double dir[3], double fov, double aspect,
                        double cam_out[12]);

/* cam = {pos, topleft, right, down}; light_p = 4 quad vertices; mode = ORC_MODE_*.
 * Builds the BVH with Bvh.Build semantics (BvhNode.fs:24-61, stable sort). */
OrcScene *orc_scene_create(const OrcPrim *prims, int n_prims,
                           const OrcMaterial *mats, int n_mats,
                           const double light_p[12], const double light_n[3],
                           const double light_color[3],
                           const double cam[12], int width, int height,
                           int max_depth, int mode);
void orc_scene_destroy(OrcScene *s);
int  orc_scene_node_slots(const OrcScene *s);             /* 2N-1                         */
void orc_scene_get_bvh(const OrcScene *s, OrcNode *nodes, int32_t *indices);
/* Replace the built tree by an externally supplied one (same layout). */
void orc_scene_set_bvh(OrcScene *s, const OrcNode *nodes, const int32_t *indices);

/* Bvh.Hit (BvhNode.fs:62-83) for n rays; prim = original primitive index or -1;
 * sub = 0/1 = which triangle of a Rect. point/normal may be NULL. */
void orc_bvh_hit(const OrcScene *s, int n, const double *origins, const double *dirs,
                 double tmin, double tmax, int32_t *prim, int32_t *sub, double *t,
                 double *point, double *normal);

/* cam.GetRay(u,v) then bvh.Hit(ray,1e-6,99999999.) (Camera.fs:134-139, Integrators.fs:108).
 * uv == NULL: pixel centres, n must be width*height, ray r = j*width+i, u=(i+.5)/w, v=(j+.5)/h. */
void orc_trace_primary(const OrcScene *s, int n, const double *uv, int32_t *prim, double *t);

/* PixelIntegrator.Sample(n) (Integrators.fs:160-172) with the counter-based RNG of
 * DESIGN.md; sample indices first_sample .. first_sample+n-1.  texture = Color[w,h],
 * element [x,y] at (x*h+y)*4, r,g,b,a doubles.  x0,x1,y0,y1 restrict the pixels
 * rendered (others are left untouched).  threads<=0: all OpenMP threads. */
void orc_sample(const OrcScene *s, int n, uint64_t seed, int first_sample,
                int x0, int y0, int x1, int y1, int threads,
                double *texture, OrcStats *stats_or_null, int count_ordered);

/* IPathTracer.TraceRay for one explicit (pixel,sample) -- the finer seam. rgb_out[3]. */
void orc_trace_path(const OrcScene *s, int px, int py, int sample, uint64_t seed, double *rgb_out);

/* Film.AddSample (Film.fs:18-23): sum += frame; target = sum / frame_count. */
void orc_film_add_sample(double *sum, const double *frame, double *target, int n_pixels,
                         double frame_count);

/* ACESFilmToneMapping + PostProcessAndToScreenBuffer (Scene.fs:273-289,315-330).
 * texture is Color[w,h] x-major; rgba8 is row-major x*4 + y*w*4. */
void orc_tonemap_rgba8(const double *texture, int width, int height, uint8_t *rgba8);

/* The RNG: Philox4x32-10, counter (c0..c3), key (k0,k1). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

int orc_max_threads(void);

/* ---- the sphere sample's integrator (mafrix_oracle_sky.c): GetColor, RenderTest/Sample/RayTracing.fs:367-382 ---- */
/* RayTraceCamera, derived (RayTracing.fs:335-358) */
typedef struct {
    double origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
    double lens_radius;
} OrcLensCamera;
typedef struct OrcSkyScene OrcSkyScene;
void orc_camera_lens(const double lookfrom[3], const double lookat[3], const double vup[3], double vfov,
                     double aspect, double aperture, double focus_dist, OrcLensCamera *out);
/* spheres: OrcPrim of kind ORC_SPHERE.  Materials: ORC_LAMBERT (ConstantTexture albedo), ORC_METAL (albedo, fuzz),
 * ORC_DIELECTRIC (ri = ei), ORC_LAMBERT_CHECKER (even = albedo, odd = (fuzz, ei, et)), ORC_LAMBERT_NOISE.
 * ranfloat[256] / perm[3*256] = Perlin's tables (may be NULL without a noise material).  max_depth = the 50 of :373. */
OrcSkyScene *orc_sky_create(const OrcPrim *spheres, int n, const OrcMaterial *mats, int n_mats,
                            const OrcLensCamera *cam, const double *ranfloat, const int32_t *perm,
                            int width, int height, int max_depth);
void orc_sky_destroy(OrcSkyScene *s);
/* ListHit(items, Ray(origin, dir), tmin, tmax), RayTracing.fs:256-258 (the Ray constructor normalises dir) */
void orc_sky_list_hit(const OrcSkyScene *s, int n, const double *origins, const double *dirs, double tmin, double tmax,
                      int32_t *prim, double *t);
/* cam.GetRay(u, v) without a lens sample + ListHit(ray, 0.00001, 10000000); uv == NULL: pixel centres */
void orc_sky_trace_primary(const OrcSkyScene *s, int n, const double *uv, int32_t *prim, double *t);
void orc_sky_trace_path(const OrcSkyScene *s, int px, int py, int sample, uint64_t seed, double *rgb_out, uint64_t *rays);
/* the pixel loop of DoRayTrace (RayTracing.fs:444-455) -> Color[w,h], element [i,j] at (i*h+j)*4 */
void orc_sky_sample(const OrcSkyScene *s, int n, uint64_t seed, int first_sample, int threads, double *texture,
                    uint64_t *rays_out);
/* sqrt + int(255.99 c) + vertical flip (RayTracing.fs:456-460): Color[w,h] -> RGBA8 row-major */
void orc_sky_display_rgba8(const double *texture, int width, int height, uint8_t *rgba8);

#ifdef __cplusplus
}
#endif
#endif
