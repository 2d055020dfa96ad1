// mfx_exact.cu -- MFX_EXACT_F64 kernels.  COMPILED WITH --fmad=false.
//
// f64 arithmetic in the operation order of the F# reference, IEEE
// divide and sqrt, no FMA contraction: primitive ids, hit distances and radiance are
// bit-identical to the reference algorithm (as restated by oracle/).  Only the traversal
// ORDER differs from Bvh.CheckHit (BvhNode.fs:62-82): near child first with a conservative
// t-shrink cull, and the reference's tie rules applied explicitly -- DESIGN.md "Exact traversal"
// argues the equivalence.
#include "mfx_exact_dev.cuh"

// ---------------------------------------------------------------- kernels
#define GRID_STRIDE_WARP(i, n) \
    for (long long i##_base = (long long)blockIdx.x * blockDim.x; i##_base < (n); i##_base += (long long)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(128) k_x_raygen(SceneX sc, WaveX w, TileMap tm, int pix0, int npix, int s0, int S, uint64_t seed)
{
    const long long total = (long long)npix * S;
    for (long long pid = (long long)blockIdx.x * blockDim.x + threadIdx.x; pid < total; pid += (long long)gridDim.x * blockDim.x) {
        const int sl = (int)(pid / npix), pl = (int)(pid - (long long)sl * npix);
        int pix, px, py;
        pixel_of(tm, sc.width, pix0 + pl, pix, px, py);
        RngX g; g.pixel = (uint32_t)pix; g.sample = (uint32_t)(s0 + sl); g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
        double u4[4];
        rng_draw_x(g, MFX_DIM_CAMERA, 0, u4);
        const double u = ((double)px + u4[0]) / (double)sc.width;      // Integrators.fs:167
        const double v = ((double)py + u4[1]) / (double)sc.height;     // Integrators.fs:168
        D3 d, org = ld3(sc.cam.pos);
        if (sc.mode == MFX_MODE_SKY) lens_ray_x(sc.cam, sc.lens, u, v, &g, org, d);     // RayTracing.fs:450-452
        else d = camera_ray_dir_x(sc.cam, u, v);
        const size_t P = (size_t)w.P;
        w.ray_o[pid] = org.x; w.ray_o[P + pid] = org.y; w.ray_o[2 * P + pid] = org.z;
        w.ray_d[pid] = d.x; w.ray_d[P + pid] = d.y; w.ray_d[2 * P + pid] = d.z;
        w.nv[pid] = 0;
        w.queue[0][pid] = (int)pid;
        if (pid == 0) w.counts[0] = (int)total;
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_x_extend(SceneX sc, WaveX w, int bounce, TravCounters *ctr)
{
    const int n = w.counts[bounce];
    const int *q = w.queue[bounce & 1];
    const size_t P = (size_t)w.P;
    unsigned long long local[3] = { 0, 0, 0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pid = q[i];
        const D3 o = mk3<double>(w.ray_o[pid], w.ray_o[P + pid], w.ray_o[2 * P + pid]);
        const D3 d = mk3<double>(w.ray_d[pid], w.ray_d[P + pid], w.ray_d[2 * P + pid]);
        const HitX h = (sc.mode == MFX_MODE_SKY) ? sky_hit_x<COUNT>(sc, o, d, MFX_SKY_TMIN, MFX_SKY_TMAX, local)     // RayTracing.fs:368
                                                 : bvh_hit_x<false, COUNT>(sc, o, d, 1e-6, 99999999., local);   // Integrators.fs:108
        w.hit_t[pid] = h.t;
        w.hit_slot[pid] = (h.slot < 0) ? -1 : (h.slot | (h.sub << 30));
    }
    if (COUNT) { for (int k = 0; k < 3; k++) if (local[k]) atomicAdd(&ctr->v[0][k], local[k]); }
}

// One path vertex: PathIntegrator.TraceRay body (Integrators.fs:109-136, mode 0) or
// NewPathTracer.TraceRay body (PathTracer.fs:25-41, mode 1), minus the two bvh.Hit calls.
__global__ void __launch_bounds__(128) k_x_shade(SceneX sc, WaveX w, TileMap tm, int pix0, int npix, int s0, int bounce, uint64_t seed)
{
    const int n = w.counts[bounce];
    const int *qin = w.queue[bounce & 1];
    int *qout = w.queue[(bounce + 1) & 1];
    const size_t P = (size_t)w.P;
    const int nwarp_iters = (n + 31) / 32;
    const int k = bounce;
    for (int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < nwarp_iters; it += gridDim.x * (blockDim.x >> 5)) {
        const int i = it * 32 + (threadIdx.x & 31);
        bool alive = false;
        int pid = -1;
        if (i < n) {
            pid = qin[i];
            const int hs = w.hit_slot[pid];
            if (hs >= 0) {
                alive = true;
                const int slot = hs & 0x3fffffff, sub = hs >> 30;
                const double t = w.hit_t[pid];
                const D3 o = mk3<double>(w.ray_o[pid], w.ray_o[P + pid], w.ray_o[2 * P + pid]);
                const D3 d = mk3<double>(w.ray_d[pid], w.ray_d[P + pid], w.ray_d[2 * P + pid]);
                const PrimX p = sc.prims[slot];
                const D3 point = o + d * t;                                   // Ray.PointAtParameter, Ray.fs:8-10
                D3 normal;
                if (p.kind == 2) normal = normalize_x(point - ld3(p.v0));     // Sphere.fs:33
                else if (sub == 0) normal = tri_normal_x(ld3(p.e1), ld3(p.e2));
                else normal = tri_normal_x(ld3(p.e2), ld3(p.e3));
                const MatX m = sc.mats[p.material];
                const int sl = pid / npix, pl = pid - sl * npix;
                int pix, px, py;
                pixel_of(tm, sc.width, pix0 + pl, pix, px, py);
                RngX g; g.pixel = (uint32_t)pix; g.sample = (uint32_t)(s0 + sl); g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);

                const double INVPI = 1. / 3.14159265358979323846, TWOPI = 2. * 3.14159265358979323846;
                D3 wi; double cr, cg, cb; double ei_shade = 0.;
                if (sc.mode == 0) {
                    // GetBxdf(): LambertianBrdf(a) for Lambertian/Metal, LambertianBrdf(Color()) for
                    // SpecularTransmission (Material.fs:52,68,121); SampleF (Material.fs:33-36)
                    const double ar = (m.kind == 2) ? 0. : m.albedo[0], ag = (m.kind == 2) ? 0. : m.albedo[1], ab = (m.kind == 2) ? 0. : m.albedo[2];
                    wi = normalize_x(random_in_unit_sphere_x(normal, g, MFX_DIM_BSDF(k)));
                    const double ei = dot(normal, wi);
                    cr = TWOPI * (ei * (INVPI * ar)); cg = TWOPI * (ei * (INVPI * ag)); cb = TWOPI * (ei * (INVPI * ab));
                } else if (m.kind == 0) {                                      // Lambertian.Scatter, Material.fs:41-46
                    wi = normalize_x(random_in_unit_sphere_x(normal, g, MFX_DIM_BSDF(k)));
                    cr = INVPI * m.albedo[0]; cg = INVPI * m.albedo[1]; cb = INVPI * m.albedo[2];
                    ei_shade = dot(normal, wi);                                // Lambertian.Shade, Material.fs:48
                } else if (m.kind == 1) {                                      // Metal.Scatter, Material.fs:61-65
                    const double fuzz = (m.fuzz < 1.0) ? m.fuzz : 1.0;
                    const D3 reflected = reflect_x(d, normal);
                    const D3 rnd = random_in_unit_sphere_x(normal, g, MFX_DIM_BSDF(k));
                    wi = normalize_x(reflected + rnd * fuzz);
                    cr = m.albedo[0]; cg = m.albedo[1]; cb = m.albedo[2];
                } else {                                                       // SpecularTransmission.Scatter, Material.fs:103-118
                    const D3 dir = -d;
                    const double cosi = dot(dir, normal);
                    double ei, et;
                    if (cosi > 0.) { ei = m.ei; et = m.et; } else { ei = m.et; et = m.ei; }
                    // Refract (Material.fs:17-24)
                    const double r = ei / et;
                    const D3 uv = normalize_x(dir);
                    const double dt = dot(uv, normal);
                    const double disc = 1.0 - r * r * (1.0 - dt * dt);
                    if (disc > 0) {
                        wi = (dir - normal * dt) * r - normal * sqrt(disc);
                        const double F = fresnel_x(m.ei, m.et, cosi);
                        const double f = (et * et) / (ei * ei);
                        const double den = fabs(dot(wi, normal));
                        cr = ((f * (1. - F)) * m.albedo[0]) / den;
                        cg = ((f * (1. - F)) * m.albedo[1]) / den;
                        cb = ((f * (1. - F)) * m.albedo[2]) / den;
                    } else {
                        wi = reflect_x(dir, normal);
                        cr = cg = cb = 0.;
                    }
                }
                // light sample: Rect.SamplePoint (Rect.fs:33-38) -> NewAreaLight.GetDirection (Light.fs:42-47)
                double u[4];
                rng_draw_x(g, MFX_DIM_LIGHT(k), 0, u);
                const D3 lp = (u[0] < 0.5) ? tri_sample_x(sc.light.t1, u[1], u[2]) : tri_sample_x(sc.light.t2, u[1], u[2]);
                const D3 toLight = lp - point;
                const double dist = sqrt(len2(toLight));
                const D3 unit = toLight / dist;
                // NewAreaLight.L (Light.fs:48-56) and `unitToLight.Dot(hit.normal) * l` (Integrators.fs:52)
                double lr = 0., lg = 0., lb = 0.;
                const double cos_o = dot(toLight, ld3(sc.light.normal));
                if (cos_o < 0.) {
                    const double solid = fabs(cos_o) * sc.light.area / len2(toLight);
                    lr = solid * sc.light.color[0]; lg = solid * sc.light.color[1]; lb = solid * sc.light.color[2];
                }
                const double dn = dot(unit, normal);
                lr = dn * lr; lg = dn * lg; lb = dn * lb;

                const size_t vb = (size_t)k * 3 * P;
                w.v_l[vb + pid] = lr; w.v_l[vb + P + pid] = lg; w.v_l[vb + 2 * P + pid] = lb;
                w.v_col[vb + pid] = cr; w.v_col[vb + P + pid] = cg; w.v_col[vb + 2 * P + pid] = cb;
                w.v_ei[(size_t)k * P + pid] = ei_shade;
                w.v_kind[(size_t)k * P + pid] = m.kind;
                w.nv[pid] = k + 1;
                w.sh_d[pid] = unit.x; w.sh_d[P + pid] = unit.y; w.sh_d[2 * P + pid] = unit.z;
                w.sh_dist[pid] = dist;
                w.ray_o[pid] = point.x; w.ray_o[P + pid] = point.y; w.ray_o[2 * P + pid] = point.z;
                w.ray_d[pid] = wi.x; w.ray_d[P + pid] = wi.y; w.ray_d[2 * P + pid] = wi.z;
            }
        }
        const int pos = warp_append(alive, &w.counts[bounce + 1]);
        if (alive) qout[pos] = pid;
    }
}

// Texture.Value(0, 0, p): ConstantTexture (RayTracing.fs:50-52), CheckerTexture (:54-61), NoiseTexture over
// Perlin.Noise (:86-99).  Checker: even = albedo, odd = (fuzz, ei, et) -- see MfxMaterial.
__device__ __forceinline__ D3 texture_value_x(const SceneX &sc, const MatX &m, D3 p)
{
    if (m.kind == 4) {
        const double sines = sin(10. * p.x) * sin(10. * p.y) * sin(10. * p.z);
        if (sines < 0.) return mk3<double>(m.fuzz, m.ei, m.et);
        return ld3(m.albedo);
    }
    if (m.kind == 5) {
        const int i = (int)(4. * p.x) & 255, j = (int)(4. * p.y) & 255, k = (int)(4. * p.z) & 255;
        const double nz = sc.perlin_rf[sc.perlin_perm[i] ^ sc.perlin_perm[256 + j] ^ sc.perlin_perm[512 + k]];
        return mk3<double>(1., 1., 1.) * nz;
    }
    return ld3(m.albedo);
}

// One level of GetColor (RayTracing.fs:367-382) minus its ListHit: miss -> sky gradient, hit -> Material.Scatter
// (Lambertian :282-290, Metal :291-299, Dielectric :300-325) and either the next ray or black.  The attenuation of
// vertex k is stored; k_x_resolve multiplies them inside-out like the recursion returns.
__global__ void __launch_bounds__(128) k_x_shade_sky(SceneX sc, WaveX w, TileMap tm, int pix0, int npix, int s0, int bounce, uint64_t seed)
{
    const int n = w.counts[bounce];
    const int *qin = w.queue[bounce & 1];
    int *qout = w.queue[(bounce + 1) & 1];
    const size_t P = (size_t)w.P;
    const int nwarp_iters = (n + 31) / 32;
    const int k = bounce;
    for (int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < nwarp_iters; it += gridDim.x * (blockDim.x >> 5)) {
        const int i = it * 32 + (threadIdx.x & 31);
        bool alive = false;
        int pid = -1;
        if (i < n) {
            pid = qin[i];
            const int hs = w.hit_slot[pid];
            const D3 o = mk3<double>(w.ray_o[pid], w.ray_o[P + pid], w.ray_o[2 * P + pid]);
            const D3 d = mk3<double>(w.ray_d[pid], w.ray_d[P + pid], w.ray_d[2 * P + pid]);
            D3 term = mk3<double>(0., 0., 0.);
            if (hs < 0) {
                const D3 unit = normalize_x(d);                               // :378-381
                const double t = 0.5 * (unit.y + 1.0);
                term = mk3<double>(1., 1., 1.) * (1.0 - t) + mk3<double>(0.5, 0.7, 1.0) * t;
            } else {
                const double t = w.hit_t[pid];
                const PrimX p = sc.prims[hs];
                const D3 point = o + d * t;                                   // PointAtParameter, :21
                const D3 normal = (point - ld3(p.v0)) / p.e1[0];              // :198
                const MatX m = sc.mats[p.material];
                const int sl = pid / npix, pl = pid - sl * npix;
                int pix, px, py;
                pixel_of(tm, sc.width, pix0 + pl, pix, px, py);
                RngX g; g.pixel = (uint32_t)pix; g.sample = (uint32_t)(s0 + sl); g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
                D3 att, wi; bool ok = true;
                if (m.kind == 1) {                                            // Metal.Scatter
                    const double fuzz = (m.fuzz < 1.0) ? m.fuzz : 1.0;
                    const D3 reflected = reflect_x(normalize_x(d), normal);
                    wi = normalize_x(reflected + random_in_unit_ball_x(g, MFX_DIM_BSDF(k)) * fuzz);
                    att = ld3(m.albedo);
                    ok = dot(wi, normal) > 0;
                } else if (m.kind == 3) {                                     // Dielectric.Scatter
                    const double ref_idx = m.ei;
                    const D3 reflected = reflect_x(d, normal);
                    D3 outward; double ni_over_nt, cosine;
                    const double dn = dot(d, normal);
                    if (dn > 0) { outward = -normal; ni_over_nt = ref_idx; cosine = ref_idx * dn; }
                    else { outward = normal; ni_over_nt = 1.0 / ref_idx; cosine = -dn; }
                    // Refract (:269-276)
                    const D3 uv = normalize_x(d);
                    const double dt = dot(uv, outward);
                    const double disc = 1.0 - ni_over_nt * ni_over_nt * (1.0 - dt * dt);
                    double reflect_prob = 1.0;
                    D3 ref_dir = mk3<double>(0., 0., 0.);
                    if (disc > 0) {
                        ref_dir = (d - outward * dt) * ni_over_nt - outward * sqrt(disc);
                        // Schlick (:277-280); (1-cosine)^5 by products, see the oracle's header
                        const double r0 = (1. - ref_idx) / (1. + ref_idx);
                        const double r1 = r0 * r0;
                        const double x = 1. - cosine, x2 = x * x, x4 = x2 * x2;
                        reflect_prob = r1 + (1. - r1) * (x4 * x);
                    }
                    double u[4];
                    rng_draw_x(g, MFX_DIM_LIGHT(k), 0, u);
                    wi = normalize_x((u[0] < reflect_prob) ? reflected : ref_dir);
                    att = mk3<double>(1., 1., 1.);
                } else {                                                      // Lambertian.Scatter over a texture
                    wi = normalize_x(normalize_x(normal) + random_in_unit_ball_x(g, MFX_DIM_BSDF(k)));
                    att = texture_value_x(sc, m, point);
                }
                if (k < sc.max_depth && ok) {                                 // `if depth < 50 && ishit`, :373
                    alive = true;
                    const size_t vb = (size_t)k * 3 * P;
                    w.v_col[vb + pid] = att.x; w.v_col[vb + P + pid] = att.y; w.v_col[vb + 2 * P + pid] = att.z;
                    w.nv[pid] = k + 1;
                    w.ray_o[pid] = point.x; w.ray_o[P + pid] = point.y; w.ray_o[2 * P + pid] = point.z;
                    w.ray_d[pid] = wi.x; w.ray_d[P + pid] = wi.y; w.ray_d[2 * P + pid] = wi.z;
                }
            }
            if (!alive) { w.v_l[pid] = term.x; w.v_l[P + pid] = term.y; w.v_l[2 * P + pid] = term.z; }
        }
        const int pos = warp_append(alive, &w.counts[bounce + 1]);
        if (alive) qout[pos] = pid;
    }
}

// Shadow query of SingleDirectLightIntegrator (Integrators.fs:44): occluded -> Color().
template <bool COUNT>
__global__ void __launch_bounds__(128) k_x_shadow(SceneX sc, WaveX w, int bounce, TravCounters *ctr)
{
    // runs over the paths that were shaded at vertex `bounce` = queue[bounce+1]
    const int n = w.counts[bounce + 1];
    const int *q = w.queue[(bounce + 1) & 1];
    const size_t P = (size_t)w.P;
    unsigned long long local[3] = { 0, 0, 0 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pid = q[i];
        const D3 o = mk3<double>(w.ray_o[pid], w.ray_o[P + pid], w.ray_o[2 * P + pid]);   // hit.point
        const D3 d = mk3<double>(w.sh_d[pid], w.sh_d[P + pid], w.sh_d[2 * P + pid]);
        const double dist = w.sh_dist[pid];
        const HitX h = bvh_hit_x<true, COUNT>(sc, o, d, 1e-6, dist - 1e-6, local);
        if (h.slot >= 0) {
            const size_t vb = (size_t)bounce * 3 * P;
            w.v_l[vb + pid] = 0.; w.v_l[vb + P + pid] = 0.; w.v_l[vb + 2 * P + pid] = 0.;
        }
    }
    if (COUNT) { for (int k = 0; k < 3; k++) if (local[k]) atomicAdd(&ctr->v[1][k], local[k]); }
}

// Unwinds the recursion inside-out so the rounding sequence equals the reference's:
//   mode 0: L_k = ((l_k / pdf_li + L_{k+1}) * col_k) / pdf          (Integrators.fs:136)
//   mode 1: L_k = l_k * col_k + col_k * Shade(L_{k+1})              (PathTracer.fs:40-41)
//   sky   : L_k = L_{k+1} * attenuation_k, L_nv = sky or black        (RayTracing.fs:374-381)
// then color <- color + L_0 per sample in sample order                (Integrators.fs:170).
__global__ void __launch_bounds__(128) k_x_resolve(SceneX sc, WaveX w, TileMap tm, int pix0, int npix, int S, double *pixsum)
{
    const size_t P = (size_t)w.P;
    const double pdf_li = 1. / sc.light.area;      // Light.fs:59
    const double PI = 3.14159265358979323846;
    for (int pl = blockIdx.x * blockDim.x + threadIdx.x; pl < npix; pl += gridDim.x * blockDim.x) {
        int pix, px, py;
        pixel_of(tm, sc.width, pix0 + pl, pix, px, py);
        double cr = pixsum[4 * (size_t)pix], cg = pixsum[4 * (size_t)pix + 1], cb = pixsum[4 * (size_t)pix + 2];
        for (int sl = 0; sl < S; sl++) {
            const size_t pid = (size_t)sl * npix + pl;
            const int nv = w.nv[pid];
            double Lr = 0., Lg = 0., Lb = 0.;
            if (sc.mode == MFX_MODE_SKY) {      // Color(c.r*attenuation.x, ..) on the way out of the recursion, RayTracing.fs:375
                Lr = w.v_l[pid]; Lg = w.v_l[P + pid]; Lb = w.v_l[2 * P + pid];
                for (int k = nv - 1; k >= 0; k--) {
                    const size_t vb = (size_t)k * 3 * P;
                    Lr = Lr * w.v_col[vb + pid]; Lg = Lg * w.v_col[vb + P + pid]; Lb = Lb * w.v_col[vb + 2 * P + pid];
                }
            } else
            for (int k = nv - 1; k >= 0; k--) {
                const size_t vb = (size_t)k * 3 * P;
                const double lr = w.v_l[vb + pid], lg = w.v_l[vb + P + pid], lb = w.v_l[vb + 2 * P + pid];
                const double kr = w.v_col[vb + pid], kg = w.v_col[vb + P + pid], kb = w.v_col[vb + 2 * P + pid];
                if (sc.mode == 0) {
                    Lr = ((lr / pdf_li + Lr) * kr) / 1.;
                    Lg = ((lg / pdf_li + Lg) * kg) / 1.;
                    Lb = ((lb / pdf_li + Lb) * kb) / 1.;
                } else {
                    double sr = Lr, sg = Lg, sb = Lb;
                    if (w.v_kind[(size_t)k * P + pid] == 0) {                  // Lambertian.Shade, Material.fs:47-50
                        const double ei = w.v_ei[(size_t)k * P + pid];
                        sr = ei * (PI * (2. * Lr)); sg = ei * (PI * (2. * Lg)); sb = ei * (PI * (2. * Lb));
                    }
                    Lr = lr * kr + kr * sr;
                    Lg = lg * kg + kg * sg;
                    Lb = lb * kb + kb * sb;
                }
            }
            cr = cr + Lr; cg = cg + Lg; cb = cb + Lb;
        }
        pixsum[4 * (size_t)pix] = cr; pixsum[4 * (size_t)pix + 1] = cg; pixsum[4 * (size_t)pix + 2] = cb;
        pixsum[4 * (size_t)pix + 3] = 1.0;
    }
}

__global__ void __launch_bounds__(128) k_x_bvh_hit(SceneX sc, int any_hit, long long n, const double *o, const double *d,
                                                   double tmin, double tmax, int *prim, int *sub, double *t)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const D3 oo = mk3<double>(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
        const D3 dd = mk3<double>(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        // MFX_SKY_TRACER: ListHit(items, Ray(origin, dir), tmin, tmax) -- the Ray constructor normalises
        const HitX h = (sc.mode == MFX_MODE_SKY) ? sky_hit_x<false>(sc, oo, normalize_x(dd), tmin, tmax, nullptr)
                     : any_hit ? bvh_hit_x<true, false>(sc, oo, dd, tmin, tmax, nullptr)
                               : bvh_hit_x<false, false>(sc, oo, dd, tmin, tmax, nullptr);
        prim[i] = (h.slot < 0) ? -1 : sc.ref_id[h.slot];
        if (sub) sub[i] = (h.slot < 0) ? 0 : h.sub;
        t[i] = (h.slot < 0) ? 0. : h.t;
    }
}

__global__ void __launch_bounds__(128) k_x_primary(SceneX sc, long long n, const double *uv, int *prim, double *t)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        double u, v;
        if (uv) { u = uv[2 * r]; v = uv[2 * r + 1]; }
        else {
            const int j = (int)(r / sc.width), i = (int)(r - (long long)j * sc.width);
            u = ((double)i + 0.5) / (double)sc.width;
            v = ((double)j + 0.5) / (double)sc.height;
        }
        HitX h;
        if (sc.mode == MFX_MODE_SKY) {      // GetRay without a lens sample + ListHit(ray, 0.00001, 10000000)
            D3 org, d;
            lens_ray_x(sc.cam, sc.lens, u, v, nullptr, org, d);
            h = sky_hit_x<false>(sc, org, d, MFX_SKY_TMIN, MFX_SKY_TMAX, nullptr);
        } else {
            const D3 d = camera_ray_dir_x(sc.cam, u, v);
            h = bvh_hit_x<false, false>(sc, ld3(sc.cam.pos), d, 1e-6, 99999999., nullptr);
        }
        prim[r] = (h.slot < 0) ? -1 : sc.ref_id[h.slot];
        t[r] = (h.slot < 0) ? 0. : h.t;
    }
}

// texture[i,j] <- color / float n (Integrators.fs:171); also the f32 row-major view.
__global__ void __launch_bounds__(256) k_finalize(const double *pixsum, int width, int height, int spp, TileMap tm,
                                                  double *color_wh, float4 *rgba_f32)
{
    const int n = tm.n_pix;
    const double dn = (double)spp;
    for (int pl = blockIdx.x * blockDim.x + threadIdx.x; pl < n; pl += gridDim.x * blockDim.x) {
        int pix, px, py;
        pixel_of(tm, width, pl, pix, px, py);
        const double r = pixsum[4 * (size_t)pix] / dn, g = pixsum[4 * (size_t)pix + 1] / dn, b = pixsum[4 * (size_t)pix + 2] / dn;
        if (color_wh) {
            double *o = color_wh + ((size_t)px * height + py) * 4;
            o[0] = r; o[1] = g; o[2] = b; o[3] = 1.0;
        }
        if (rgba_f32) rgba_f32[pix] = make_float4((float)r, (float)g, (float)b, 1.0f);
    }
}

// ---------------------------------------------------------------- launchers
void mfx_x_raygen(const LaunchCfg &c, const SceneX &sc, const WaveX &w, TileMap tm, int pix0, int npix, int s0, int S, uint64_t seed)
{
    k_x_raygen<<<persistent_blocks(k_x_raygen, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, tm, pix0, npix, s0, S, seed);
}
void mfx_x_extend(const LaunchCfg &c, const SceneX &sc, const WaveX &w, int bounce, TravCounters *ctr)
{
    if (ctr) k_x_extend<true><<<persistent_blocks(k_x_extend<true>, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, bounce, ctr);
    else k_x_extend<false><<<persistent_blocks(k_x_extend<false>, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, bounce, ctr);
}
void mfx_x_shade(const LaunchCfg &c, const SceneX &sc, const WaveX &w, TileMap tm, int pix0, int npix, int s0, int bounce, uint64_t seed)
{
    k_x_shade<<<persistent_blocks(k_x_shade, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, tm, pix0, npix, s0, bounce, seed);
}
void mfx_x_shade_sky(const LaunchCfg &c, const SceneX &sc, const WaveX &w, TileMap tm, int pix0, int npix, int s0, int bounce, uint64_t seed)
{
    k_x_shade_sky<<<persistent_blocks(k_x_shade_sky, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, tm, pix0, npix, s0, bounce, seed);
}
void mfx_x_shadow(const LaunchCfg &c, const SceneX &sc, const WaveX &w, int bounce, TravCounters *ctr)
{
    if (ctr) k_x_shadow<true><<<persistent_blocks(k_x_shadow<true>, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, bounce, ctr);
    else k_x_shadow<false><<<persistent_blocks(k_x_shadow<false>, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, bounce, ctr);
}
void mfx_x_resolve(const LaunchCfg &c, const SceneX &sc, const WaveX &w, TileMap tm, int pix0, int npix, int S, double *pixsum)
{
    k_x_resolve<<<persistent_blocks(k_x_resolve, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, w, tm, pix0, npix, S, pixsum);
}
void mfx_x_bvh_hit(const LaunchCfg &c, const SceneX &sc, int any_hit, long long n, const double *o, const double *d,
                   double tmin, double tmax, int *prim, int *sub, double *t)
{
    k_x_bvh_hit<<<persistent_blocks(k_x_bvh_hit, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, any_hit, n, o, d, tmin, tmax, prim, sub, t);
}
void mfx_x_primary(const LaunchCfg &c, const SceneX &sc, long long n, const double *uv, int *prim, double *t)
{
    k_x_primary<<<persistent_blocks(k_x_primary, c.threads, c.blocks), c.threads, 0, c.stream>>>(sc, n, uv, prim, t);
}
void mfx_x_finalize(const LaunchCfg &c, const double *pixsum, int width, int height, double, int spp, TileMap tm,
                    double *color_wh, float4 *rgba_f32)
{
    k_finalize<<<persistent_blocks(k_finalize, 256, c.blocks), 256, 0, c.stream>>>(pixsum, width, height, spp, tm, color_wh, rgba_f32);
}
