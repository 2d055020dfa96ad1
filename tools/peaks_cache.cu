// peaks_cache.cu -- measured on-chip read bandwidth of THIS GPU for the access pattern of the traversal kernels, the
// denominator of bench.py's roofline for L2-resident scenes (MEASURED_PEAKS.json only has the HBM copy figure).
//
// k_f_trace6 fetches one 128-byte four-child record per lane and node step as four 256-bit loads (LDG.E.256,
// ld.global.nc.v8.f32) -- every lane of a divergent warp at its own record -- and 48-byte triangle slots as 128-bit loads.
// The L1 serves such a warp instruction sector by sector, so what bounds the kernel is not 128 B/clk/SM of coalesced
// bandwidth but the rate at which the L1 / L2 turn divergent 32-byte sectors around.  This program measures exactly that:
// persistent warps, every lane chasing its own pseudo-random sequence of 128-byte records inside a table of F bytes, the
// same four v8 loads per record, `ilp` independent records in flight per lane; plus the coalesced variant (all lanes of a
// warp in one 4 KB stretch) as the ceiling.  F sweeps 64 KB (L1 resident) .. 512 MB (HBM).  Prints one JSON line.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks_cache peaks_cache.cu && ./peaks_cache
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ float ld_record(const float4 *p)
{
    float a[8], b[8], c[8], d[8];
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]), "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7]) : "l"(p + 2));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]), "=f"(c[4]), "=f"(c[5]), "=f"(c[6]), "=f"(c[7]) : "l"(p + 4));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p + 6));
    return a[0] + a[7] + b[0] + b[7] + c[0] + c[7] + d[0] + d[7];
}

// records: table of n_rec 128-byte records.  DIVERGENT: every lane follows its own LCG through the table (independent of
// the loaded data, like a ray that already knows the link of its next record: throughput, not latency, is measured).
template <bool DIVERGENT, int ILP>
__global__ void __launch_bounds__(128) k_chase(const float4 *records, unsigned n_rec_mask, int iters, float *sink)
{
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    if (!DIVERGENT) s = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            s = s * 1664525u + 1013904223u;
            unsigned r = (s >> 7) & n_rec_mask;
            if (!DIVERGENT) r = (r & ~31u) | (threadIdx.x & 31u);        // one warp = 32 consecutive records = 4 KB
            acc += ld_record(records + (size_t)r * 8);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <bool DIVERGENT, int ILP>
static double run(const float4 *tab, size_t bytes, int sms, float *sink)
{
    const unsigned n_rec = (unsigned)(bytes / 128);
    const int blocks = sms * 8, threads = 128, iters = 2048 / ILP;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_chase<DIVERGENT, ILP><<<blocks, threads>>>(tab, n_rec - 1, iters / 4, sink);       // warm the caches
    double best = 0.;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        k_chase<DIVERGENT, ILP><<<blocks, threads>>>(tab, n_rec - 1, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double gbs = (double)blocks * threads * iters * ILP * 128.0 / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

int main()
{
    int dev = 0, sms = 0, mhz = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev));
    const size_t max_bytes = (size_t)512 << 20;
    float4 *tab = nullptr; float *sink = nullptr;
    CHECK(cudaMalloc(&tab, max_bytes)); CHECK(cudaMalloc(&sink, 4));
    CHECK(cudaMemset(tab, 0, max_bytes));
    const size_t sizes[] = { (size_t)64 << 10, (size_t)512 << 10, (size_t)4 << 20, (size_t)32 << 20, (size_t)512 << 20 };
    const char *names[] = { "64KB", "512KB", "4MB", "32MB", "512MB" };
    printf("{\"sms\": %d, \"sm_khz\": %d, \"pattern\": \"4 x ld.global.nc.v8.f32 per 128-byte record, one record per lane\", \"unit\": \"GB/s of requested bytes\"", sms, mhz);
    for (int i = 0; i < 5; i++) {
        const double d4 = run<true, 4>(tab, sizes[i], sms, sink), d1 = run<true, 1>(tab, sizes[i], sms, sink);
        const double c4 = run<false, 4>(tab, sizes[i], sms, sink);
        printf(", \"divergent_%s\": %.1f, \"divergent_ilp1_%s\": %.1f, \"coalesced_%s\": %.1f", names[i], d4, names[i], d1, names[i], c4);
    }
    printf("}\n");
    CHECK(cudaDeviceSynchronize());
    cudaFree(tab); cudaFree(sink);
    return 0;
}
