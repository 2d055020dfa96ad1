// mfx_fast_dev.cuh -- f32 device helpers shared by the wavefront kernels (mfx_fast.cu) and the id-exact hybrid
// traversal (mfx_hybrid.cu): ray setup, 256-bit record loads, the 4-key sorting helpers, queue-count indices.
#pragma once
#include "mfx_device.cuh"

typedef V3<float> F3;

#define FAST_BLOCK 128
#define FAST_LEVELS 30

__device__ __forceinline__ F3 f3(float x, float y, float z) { return mk3<float>(x, y, z); }
__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }
// One 256-bit read-only load (sm_100: LDG.E.256): two adjacent float4 of a 32-byte aligned pair.  A divergent warp
// pays the L1 per 32-byte sector touched, so fetching a record as 4 x 32 B instead of 7 x 16 B halves that cost.
__device__ __forceinline__ void ldg8(const float4 *p, float4 &a, float4 &b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

struct RayF {
    F3 o, d, idir, ood;
    float tmin;
    int src;    // leaf-order primitive the ray starts on (never re-hit a planar source), -1 none
};

__device__ __forceinline__ float safe_rcp(float d)
{
    const float eps = 1e-30f;
    return 1.0f / (fabsf(d) > eps ? d : copysignf(eps, d));
}

__device__ __forceinline__ RayF make_ray(F3 o, F3 d, float tmin, int src)
{
    RayF r; r.o = o; r.d = d; r.tmin = tmin; r.src = src;
    r.idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    r.ood = f3(o.x * r.idir.x, o.y * r.idir.y, o.z * r.idir.z);
    return r;
}

// slab test; returns entry distance, hit iff entry <= exit
__device__ __forceinline__ bool box_f(const RayF &r, float lx, float ly, float lz, float hx, float hy, float hz,
                                      float tmax, float &entry)
{
    const float x0 = fmaf(lx, r.idir.x, -r.ood.x), x1 = fmaf(hx, r.idir.x, -r.ood.x);
    const float y0 = fmaf(ly, r.idir.y, -r.ood.y), y1 = fmaf(hy, r.idir.y, -r.ood.y);
    const float z0 = fmaf(lz, r.idir.z, -r.ood.z), z1 = fmaf(hz, r.idir.z, -r.ood.z);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), r.tmin));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    entry = tn;
    return tn <= tf;
}

// Intersects the slots of one leaf: Moller-Trumbore with the reference's acceptance rules
// (Trangle.fs:130-148) folded into one predicate (no early exits: every lane of a leaf vote runs the
// same instructions), the stable quadratic of Sphere.fs:21-43 for spheres.
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ RayF make_ray_fast(F3 o, F3 d, float tmin, int src)
{
    RayF r; r.o = o; r.d = d; r.tmin = tmin; r.src = src;
    const float eps = 1e-30f;
    r.idir = f3(rcp_approx(fabsf(d.x) > eps ? d.x : copysignf(eps, d.x)),
                rcp_approx(fabsf(d.y) > eps ? d.y : copysignf(eps, d.y)),
                rcp_approx(fabsf(d.z) > eps ? d.z : copysignf(eps, d.z)));
    r.ood = f3(o.x * r.idir.x, o.y * r.idir.y, o.z * r.idir.z);
    return r;
}


#define CNT_SH(b)  (MFX_MAX_VERTS + 2 + (b))
#define CUR_EXT(b) (2 * (MFX_MAX_VERTS + 2) + (b))
#define CUR_SH(b)  (3 * (MFX_MAX_VERTS + 2) + (b))

#define KEY_INF 0x7f800000u

__device__ __forceinline__ unsigned umin_(unsigned a, unsigned b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned umax_(unsigned a, unsigned b) { return a > b ? a : b; }
__device__ __forceinline__ int pick4(const float4 &m, unsigned id)
{
    const float lo = (id & 1u) ? m.y : m.x, hi = (id & 1u) ? m.w : m.z;
    return __float_as_int((id & 2u) ? hi : lo);
}
