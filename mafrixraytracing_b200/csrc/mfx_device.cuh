// mfx_device.cuh -- device helpers shared by the exact and fast kernel TUs:
// small vector template, the counter-based RNG, warp-aggregated queue append.
#pragma once
#include "mfx_internal.h"

template <typename T>
struct V3 {
    T x, y, z;
};
template <typename T> __device__ __forceinline__ V3<T> mk3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __device__ __forceinline__ V3<T> ld3(const T *p) { return mk3<T>(p[0], p[1], p[2]); }
// Operators in the shapes Core/Point.fs defines them (one rounding per written operation).
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a) { return mk3<T>(-a.x, -a.y, -a.z); }
template <typename T> __device__ __forceinline__ V3<T> operator*(V3<T> v, T a) { return mk3<T>(v.x * a, v.y * a, v.z * a); }
template <typename T> __device__ __forceinline__ V3<T> operator/(V3<T> v, T a) { return mk3<T>(v.x / a, v.y / a, v.z / a); }
template <typename T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> v)
{
    return mk3<T>(a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x);
}
template <typename T> __device__ __forceinline__ T len2(V3<T> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// counter = (pixel, sample, dim, iter), key = seed.  See DESIGN.md "RNG".
__device__ __forceinline__ void philox4x32_10(uint32_t pixel, uint32_t sample, uint32_t dim, uint32_t iter,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
    uint32_t c[4] = { pixel, sample, dim, iter };
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        philox_round(c, k0, k1);
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
#define MFX_DIM_CAMERA 0u
#define MFX_DIM_BSDF(k) (1u + 2u * (uint32_t)(k))
#define MFX_DIM_LIGHT(k) (2u + 2u * (uint32_t)(k))

__device__ __forceinline__ double u32_to_unit_f64(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }
// f32 uniform in [0,1): top 24 bits (differs from the f64 stream value by < 2^-24)
__device__ __forceinline__ float u32_to_unit_f32(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------- debug build (make DEBUG=1)
#ifdef MFX_DEBUG_CHECKS
#define DBG_CHECK(cond, counts) do { if (!(cond)) atomicAdd(&(counts)[MFX_DBG_SLOT], 1); } while (0)
#else
#define DBG_CHECK(cond, counts) do { } while (0)
#endif

// ---------------------------------------------------------------- warp-aggregated append
// Lanes with `pred` obtain consecutive positions in a queue with ONE atomic per warp:
// ballot -> popc -> leader atomicAdd -> shuffle broadcast.  Must be called by the full warp
// convergently (callers keep loops warp-uniform).
__device__ __forceinline__ int warp_append(bool pred, int *counter)
{
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return -1;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ void pixel_of(const TileMap &tm, int width, int pl, int &pix, int &px, int &py)
{
    if (tm.stripe > 0) {
        // every owned stripe but possibly the last one is full, so the local stripe index is a plain division
        const int per = tm.stripe * tm.height;
        const int ls = pl / per, rem = pl - ls * per;
        const int x0 = (ls * tm.world + tm.rank) * tm.stripe;
        const int wl = min(tm.stripe, width - x0);
        py = rem / wl;
        px = x0 + (rem - py * wl);
        pix = py * width + px;
        return;
    }
    pix = tm.pix ? tm.pix[pl] : pl;
    py = pix / width;
    px = pix - py * width;
}

// Path id of a wave <-> (sample, pixel): blocks of 2^sshift samples of one pixel are neighbours, pixels follow in
// pixel_of's order, then the next block of samples.  With 8-pixel stripe rows and sshift = 2 the 32 rays a warp picks up
// together are 4 samples each of 8 adjacent pixels -- the coherence of primary rays lasts into their shadow rays and
// the first bounce.  (Which path carries which sample never shows in a frame: the RNG is keyed on pixel and sample.)
__device__ __forceinline__ void path_split(const TileMap &tm, long long pid, int npix, int &sl, int &pl)
{
    const long long q = pid >> tm.sshift;
    const int slb = (int)(q / npix);
    pl = (int)(q - (long long)slb * npix);
    sl = (slb << tm.sshift) | (int)(pid & ((1 << tm.sshift) - 1));
}
__device__ __forceinline__ size_t path_join(const TileMap &tm, int sl, int pl, int npix)
{
    return ((((size_t)(sl >> tm.sshift)) * (size_t)npix + (size_t)pl) << tm.sshift) | (size_t)(sl & ((1 << tm.sshift) - 1));
}

// Persistent grids: SM count x resident blocks per SM for this kernel (cached per kernel, dynamic shared memory and
// device; guarded: distinct handles may launch from distinct host threads, one per GPU).
#include <map>
#include <mutex>
#include <tuple>
template <typename K>
static inline int persistent_blocks(K kernel, int threads, int sms, size_t dyn_smem = 0)
{
    static std::mutex mu;
    static std::map<std::tuple<const void *, size_t, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::tuple<const void *, size_t, int> key((const void *)kernel, dyn_smem, dev);
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, dyn_smem) != cudaSuccess || occ < 1) occ = 1;
        it = cache.emplace(key, occ).first;
    }
    return sms * it->second;
}
