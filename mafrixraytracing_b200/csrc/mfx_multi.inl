// mfx_multi.inl -- multi-GPU behind the C ABI (included at the end of mfx_host.cpp: same translation unit).
//
// The reference host is ONE process with ONE render thread (Film.fs:67-73 calls Scene.Render, Integrators.fs:160-172 is
// called from there); its only parallelism is Array.Parallel.iter over pixels (:164).  So the N-GPU path lives inside
// the library: mfx_multi_create replicates the scene on every listed device (the reference tree is built once, the
// own SAH tree once -- content cache of flatten_fast) and starts one worker thread per device; mfx_multi_sample, called
// from the host's single thread, hands every worker the same Sample(n) with its own stripe set and returns when the
// caller's texture is complete.
//
// Ownership is by COLUMN STRIPES (TileMap in mfx_internal.h): stripe c -> device c % N, computed arithmetically in the
// kernels.  In the reference's x-major Color[w,h] (Texture.fs:21-28) a stripe is contiguous, so each device writes its
// share straight into the caller's texture with ONE strided cudaMemcpy2DAsync over its own PCIe link -- no reduce, no
// gather, no full-frame buffer crossing NVLink: the path has no exchange step, so it gets no collective.  With the
// counter-based RNG keyed on absolute pixel / sample the assembled frame is bit-identical to the one-GPU frame.
#include <condition_variable>

struct MfxMulti {
    std::vector<int> devices;
    std::vector<MfxScene *> scenes;
    int width = 0, height = 0, stripe = 16;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t job = 0;               // bumped per call; a worker runs each job id once
    int pending = 0;
    bool quit = false;
    // the job
    MfxSampleParams params;
    double *texture = nullptr;      // Color[w,h] or
    float  *rgba = nullptr;         // row-major float4
    std::vector<int> rc;
    std::vector<std::string> err;
    std::vector<double> ms_wall;    // per device: run_sample + its D2H, host clock
    double ms_call = 0.;
    bool posted = false;            // a job was handed to the workers and not yet waited for
    bool prepare_only = false;      // ... and it is mfx_multi_prepare's (layouts only)
    std::chrono::steady_clock::time_point t_post;
};

// this device's stripes of the finished frame -> the caller's buffer
static int multi_copy_out(MfxMulti *m, int i)
{
    MfxScene *s = m->scenes[(size_t)i];
    const int N = (int)m->devices.size(), T = m->stripe, W = m->width, H = m->height;
    cudaStream_t st = s->stream;
    if (m->texture) {
        // x-major: stripe at column x0 = bytes [x0*H*32, (x0+wl)*H*32); this rank's full stripes are N*T columns apart
        const size_t col = (size_t)H * 4 * sizeof(double);
        int n_full = 0, x_part = -1;
        for (int x0 = i * T; x0 < W; x0 += N * T) { if (x0 + T <= W) n_full++; else x_part = x0; }
        const size_t off = (size_t)i * T * col;
        if (n_full > 0)
            CUDA_TRY(cudaMemcpy2DAsync((char *)m->texture + off, (size_t)N * T * col, (const char *)s->d_color_wh + off, (size_t)N * T * col,
                                       (size_t)T * col, (size_t)n_full, cudaMemcpyDeviceToHost, st));
        if (x_part >= 0)
            CUDA_TRY(cudaMemcpyAsync((char *)m->texture + (size_t)x_part * col, (const char *)s->d_color_wh + (size_t)x_part * col,
                                     (size_t)(W - x_part) * col, cudaMemcpyDeviceToHost, st));
    } else {
        for (int x0 = i * T; x0 < W; x0 += N * T) {        // row-major: a stripe is H pieces of wl float4
            const int wl = std::min(T, W - x0);
            CUDA_TRY(cudaMemcpy2DAsync(m->rgba + 4 * (size_t)x0, (size_t)W * sizeof(float4), s->d_rgba + x0, (size_t)W * sizeof(float4),
                                       (size_t)wl * sizeof(float4), (size_t)H, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return MFX_OK;
}

static void multi_worker(MfxMulti *m, int i)
{
    cudaSetDevice(m->devices[(size_t)i]);
    g_device = m->devices[(size_t)i];
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_job.wait(lk, [&] { return m->quit || m->job != seen; });
            if (m->quit) break;
            seen = m->job;
        }
        const auto t0 = std::chrono::steady_clock::now();
        MfxScene *s = m->scenes[(size_t)i];
        MfxSampleParams p = m->params;
        p.rank = i; p.world = (int)m->devices.size(); p.tile_size = m->stripe;
        p.flags |= MFX_SAMPLE_STRIPES | MFX_SAMPLE_NO_CLEAR;
        int rc = ensure_frame_buffers(s);
        if (m->prepare_only) {      // mfx_multi_prepare: the layouts of this precision, no frame
            if (rc == MFX_OK) rc = mfx_scene_prepare(s, m->params.precision);
            m->rc[(size_t)i] = rc;
            if (rc != MFX_OK) m->err[(size_t)i] = g_err;
            std::lock_guard<std::mutex> lk(m->mu);
            if (--m->pending == 0) m->cv_done.notify_all();
            continue;
        }
        if (rc == MFX_OK) rc = run_sample(s, &p, m->texture ? s->d_color_wh : nullptr, m->texture ? nullptr : s->d_rgba);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc == MFX_OK) rc = multi_copy_out(m, i);
        if (env_long("MFX_DEBUG", 0))
            fprintf(stderr, "[mfx] multi worker %d: run_sample %.2f ms (device %.2f ms), copy out %.2f ms\n", i,
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), s->stats.ms_total,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
        m->rc[(size_t)i] = rc;
        if (rc != MFX_OK) m->err[(size_t)i] = g_err;
        m->ms_wall[(size_t)i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        {
            std::lock_guard<std::mutex> lk(m->mu);
            if (--m->pending == 0) m->cv_done.notify_all();
        }
    }
    // every worker takes its own replica down (stream sync, ~60 buffers back to the pool, events): side by side
    mfx_scene_destroy(m->scenes[(size_t)i]);
    m->scenes[(size_t)i] = nullptr;
}

extern "C" int mfx_multi_create(const MfxSceneDesc *d, const int32_t *devices, int32_t n_devices, MfxMulti **out)
{
    if (!d || !out) return fail(MFX_ERR_INVALID_ARGUMENT, "mfx_multi_create: null argument");
    *out = nullptr;
    const int avail = mfx_device_count();
    if (avail <= 0) return fail(MFX_ERR_NO_DEVICE, "no CUDA device visible: libmafrix_cuda has no CPU fallback");
    std::vector<int> devs;
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else for (int i = 0; i < (n_devices > 0 ? std::min(n_devices, avail) : avail); i++) devs.push_back(i);     // NULL: the first n (or all)
    for (size_t i = 0; i < devs.size(); i++) {
        if (devs[i] < 0 || devs[i] >= avail) return fail(MFX_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", devs[i], avail - 1);
        for (size_t j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(MFX_ERR_INVALID_ARGUMENT, "device %d listed twice", devs[i]);
    }
    const int keep = g_device;
    MfxMulti *m = new MfxMulti();
    m->devices = devs;
    m->width = d->width; m->height = d->height;
    m->stripe = (int)std::max(1L, env_long("MFX_MULTI_STRIPE", 16));
    m->scenes.assign(devs.size(), nullptr);
    // replica 0 on this thread (argument checks, and Bvh.Build -- BvhNode.fs:24-61 -- runs once: the others take its tree),
    // the other replicas side by side on short-lived threads
    MfxSceneDesc dd = *d;
    int rc = mfx_init(devs[0]);
    if (rc == MFX_OK) rc = mfx_scene_create(&dd, &m->scenes[0]);
    std::string why = g_err;
    if (rc == MFX_OK) {
        m->scenes[0]->in_process_replica = true;    // the workers share one process: the tree cache lets one of them build, with every core
        if (!dd.nodes) { dd.nodes = m->scenes[0]->nodes.data(); dd.n_node_slots = (int32_t)m->scenes[0]->nodes.size(); dd.indices = m->scenes[0]->indices.data(); }
        std::vector<int> rcs(devs.size(), MFX_OK);
        std::vector<std::string> errs(devs.size());
        std::vector<std::thread> makers;
        for (size_t i = 1; i < devs.size(); i++)
            makers.emplace_back([&, i] {
                int r = mfx_init(devs[i]);
                if (r == MFX_OK) r = mfx_scene_create(&dd, &m->scenes[i]);
                if (r == MFX_OK) m->scenes[i]->in_process_replica = true;
                rcs[i] = r;
                if (r != MFX_OK) errs[i] = g_err;
            });
        for (std::thread &th : makers) th.join();
        for (size_t i = 1; i < devs.size() && rc == MFX_OK; i++) if (rcs[i] != MFX_OK) { rc = rcs[i]; why = errs[i]; }
    }
    g_device = keep;                     // the calling thread keeps the device it had chosen (or none)
    if (keep >= 0) cudaSetDevice(keep);
    if (rc != MFX_OK) {
        for (MfxScene *s : m->scenes) mfx_scene_destroy(s);
        delete m;
        g_err = why;
        return rc;
    }
    const size_t N = devs.size();
    m->rc.assign(N, MFX_OK); m->err.assign(N, std::string()); m->ms_wall.assign(N, 0.);
    for (size_t i = 0; i < N; i++) m->threads.emplace_back(multi_worker, m, (int)i);
    *out = m;
    return MFX_OK;
}

static int multi_wait(MfxMulti *m);

extern "C" int mfx_multi_destroy(MfxMulti *m)
{
    if (!m) return MFX_OK;
    multi_wait(m);                       // a frame still in flight finishes first
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cv_job.notify_all();
    for (std::thread &t : m->threads) t.join();
    for (MfxScene *s : m->scenes) mfx_scene_destroy(s);          // whatever a worker did not take down itself
    delete m;
    return MFX_OK;
}

extern "C" int mfx_multi_device_count(const MfxMulti *m, int32_t *n_out)
{
    if (!m || !n_out) return fail(MFX_ERR_INVALID_ARGUMENT, "null argument");
    *n_out = (int32_t)m->devices.size();
    return MFX_OK;
}

// Posts one Sample to the workers and returns; multi_wait completes it.  One job in flight per handle.
static int multi_post(MfxMulti *m, const MfxSampleParams *p, double *texture, float *rgba)
{
    if (!m || !p) return fail(MFX_ERR_INVALID_ARGUMENT, "null argument");
    if (p->spp <= 0) return fail(MFX_ERR_INVALID_ARGUMENT, "spp must be positive, got %d", p->spp);
    if (p->world > 1) return fail(MFX_ERR_INVALID_ARGUMENT, "mfx_multi_sample shards the frame itself: pass world <= 1");
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->pending != 0) return fail(MFX_ERR_INVALID_ARGUMENT, "a frame is still in flight on this handle: call mfx_multi_wait first");
    m->t_post = std::chrono::steady_clock::now();
    m->params = *p; m->texture = texture; m->rgba = rgba;
    m->prepare_only = (texture == nullptr && rgba == nullptr);
    m->pending = (int)m->devices.size();
    m->posted = true;
    m->job++;
    m->cv_job.notify_all();
    return MFX_OK;
}

static int multi_wait(MfxMulti *m)
{
    if (!m) return fail(MFX_ERR_INVALID_ARGUMENT, "null argument");
    {
        std::unique_lock<std::mutex> lk(m->mu);
        if (!m->posted) return MFX_OK;
        m->cv_done.wait(lk, [&] { return m->pending == 0; });
        m->posted = false;
    }
    m->ms_call = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - m->t_post).count();
    for (size_t i = 0; i < m->devices.size(); i++)
        if (m->rc[i] != MFX_OK) return fail(m->rc[i], "device %d: %s", m->devices[i], m->err[i].c_str());
    return MFX_OK;
}

extern "C" int mfx_multi_sample(MfxMulti *m, const MfxSampleParams *p, double *texture)
{
    if (!texture) return fail(MFX_ERR_INVALID_ARGUMENT, "null texture");
    MFX_TRY(multi_post(m, p, texture, nullptr));
    return multi_wait(m);
}

extern "C" int mfx_multi_sample_f32(MfxMulti *m, const MfxSampleParams *p, float *rgba)
{
    if (!rgba) return fail(MFX_ERR_INVALID_ARGUMENT, "null output");
    MFX_TRY(multi_post(m, p, nullptr, rgba));
    return multi_wait(m);
}

// Sample without the wait: the workers render and download while the caller goes on (e.g. builds the next frame's scene).
extern "C" int mfx_multi_sample_async(MfxMulti *m, const MfxSampleParams *p, double *texture)
{
    if (!texture) return fail(MFX_ERR_INVALID_ARGUMENT, "null texture");
    return multi_post(m, p, texture, nullptr);
}

extern "C" int mfx_multi_wait(MfxMulti *m) { return multi_wait(m); }

// Every replica builds the device layouts of this precision now (side by side) instead of inside its first Sample: a host
// that pipelines frames calls it while the previous frame renders.
extern "C" int mfx_multi_prepare(MfxMulti *m, int32_t precision)
{
    MfxSampleParams p; memset(&p, 0, sizeof(p));
    p.precision = precision; p.spp = 1; p.world = 1;
    MFX_TRY(multi_post(m, &p, nullptr, nullptr));
    return multi_wait(m);
}

// total: rays / paths / launches summed over the devices, times = the slowest device (they run side by side);
// per_device (n_devices entries, may be NULL): each device's own MfxStats.
extern "C" int mfx_multi_get_stats(const MfxMulti *m, MfxStats *total, MfxStats *per_device)
{
    if (!m || !total) return fail(MFX_ERR_INVALID_ARGUMENT, "null argument");
    memset(total, 0, sizeof(*total));
    for (size_t i = 0; i < m->scenes.size(); i++) {
        const MfxStats &a = m->scenes[i]->stats;
        if (per_device) per_device[i] = a;
        total->closest_rays += a.closest_rays; total->shadow_rays += a.shadow_rays; total->paths += a.paths;
        for (int c = 0; c < 2; c++) { total->nodes[c] += a.nodes[c]; total->tris[c] += a.tris[c]; total->spheres[c] += a.spheres[c]; }
        total->ms_total = std::max(total->ms_total, a.ms_total); total->ms_extend = std::max(total->ms_extend, a.ms_extend);
        total->ms_shadow = std::max(total->ms_shadow, a.ms_shadow); total->ms_shade = std::max(total->ms_shade, a.ms_shade);
        total->launches += a.launches; total->launches_extend += a.launches_extend; total->launches_shadow += a.launches_shadow;
        total->hybrid_fixups += a.hybrid_fixups;
    }
    return MFX_OK;
}
