// Tree-quality experiment (not part of the product): how many node / triangle tests does an ordered,
// t-shrinking traversal need per ray on (R) the reference's median-split tree (BvhNode.fs:42-61) and on
// (S) a binned-SAH tree over the same triangles?  Rays: C2-style primary rays, one uniform-hemisphere bounce
// and one shadow ray per hit.  Usage: tree_quality tris.bin  (tris.bin = n x 9 doubles), see tools/tree_quality.py
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; int left, right, first, count; } Node;   // leaf: count > 0
typedef struct { float v0[3], e1[3], e2[3]; } Tri;

static int ntri; static Tri *tris; static float (*tlo)[3], (*thi)[3], (*tc)[3];
static Node *nodes; static int nnodes; static int *idx;
static int g_axis;

static int cmp_centroid(const void *a, const void *b)
{
    float x = tc[*(const int *)a][g_axis], y = tc[*(const int *)b][g_axis];
    return (x > y) - (x < y);
}
static void bound(int first, int count, float *lo, float *hi)
{
    for (int a = 0; a < 3; a++) { lo[a] = 1e30f; hi[a] = -1e30f; }
    for (int k = 0; k < count; k++) for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], tlo[idx[first + k]][a]); hi[a] = fmaxf(hi[a], thi[idx[first + k]][a]);
    }
}
static float area(const float *lo, const float *hi)
{
    float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    return 2.f * (x * y + y * z + z * x);
}

static int build(int first, int count, int mode, int maxleaf)
{
    int me = nnodes++;
    Node *n = &nodes[me];
    bound(first, count, n->lo, n->hi);
    n->first = first; n->count = 0; n->left = n->right = -1;
    if (mode == 0) {                                   // reference: longest axis of the node box, median of count
        if (count <= 3) { nodes[me].count = count; return me; }
        float dx = n->hi[0] - n->lo[0], dy = n->hi[1] - n->lo[1], dz = n->hi[2] - n->lo[2];
        g_axis = (dx > dy && dx > dz) ? 0 : (dy > dz ? 1 : 2);
        qsort(idx + first, count, sizeof(int), cmp_centroid);
        int lc = count / 2;
        int l = build(first, lc, mode, maxleaf), r = build(first + lc, count - lc, mode, maxleaf);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    // full-sweep SAH over the three axes
    if (count == 1) { nodes[me].count = 1; return me; }
    float best = 1e30f; int baxis = -1, bsplit = -1;
    float *rarea = malloc(sizeof(float) * count);
    for (int ax = 0; ax < 3; ax++) {
        g_axis = ax;
        qsort(idx + first, count, sizeof(int), cmp_centroid);
        float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
        for (int k = count - 1; k > 0; k--) {
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], tlo[idx[first + k]][a]); hi[a] = fmaxf(hi[a], thi[idx[first + k]][a]); }
            rarea[k] = area(lo, hi);
        }
        for (int a = 0; a < 3; a++) { lo[a] = 1e30f; hi[a] = -1e30f; }
        for (int k = 1; k < count; k++) {
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], tlo[idx[first + k - 1]][a]); hi[a] = fmaxf(hi[a], thi[idx[first + k - 1]][a]); }
            float c = area(lo, hi) * k + rarea[k] * (count - k);
            if (c < best) { best = c; baxis = ax; bsplit = k; }
        }
    }
    free(rarea);
    float leafcost = area(n->lo, n->hi) * count;
    if (count <= maxleaf && leafcost <= best + 1.0f * area(n->lo, n->hi)) { nodes[me].count = count; return me; }
    g_axis = baxis;
    qsort(idx + first, count, sizeof(int), cmp_centroid);
    int l = build(first, bsplit, mode, maxleaf), r = build(first + bsplit, count - bsplit, mode, maxleaf);
    nodes[me].left = l; nodes[me].right = r;
    return me;
}

static long long c_nodes, c_tris, c_steps;
static int box(const Node *n, const float *o, const float *id, float tmax, float *entry)
{
    float tn = 1e-6f, tf = tmax;
    for (int a = 0; a < 3; a++) {
        float t0 = (n->lo[a] - o[a]) * id[a], t1 = (n->hi[a] - o[a]) * id[a];
        tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    }
    *entry = tn;
    return tn <= tf;
}
static int tri_hit(const Tri *t, const float *o, const float *d, float *tout)
{
    float s1[3] = { d[1] * t->e2[2] - d[2] * t->e2[1], d[2] * t->e2[0] - d[0] * t->e2[2], d[0] * t->e2[1] - d[1] * t->e2[0] };
    float div = s1[0] * t->e1[0] + s1[1] * t->e1[1] + s1[2] * t->e1[2];
    if (fabsf(div) < 1e-9f) return 0;
    float inv = 1.f / div, dd[3] = { o[0] - t->v0[0], o[1] - t->v0[1], o[2] - t->v0[2] };
    float b1 = (dd[0] * s1[0] + dd[1] * s1[1] + dd[2] * s1[2]) * inv;
    if (b1 < 0 || b1 > 1) return 0;
    float s2[3] = { dd[1] * t->e1[2] - dd[2] * t->e1[1], dd[2] * t->e1[0] - dd[0] * t->e1[2], dd[0] * t->e1[1] - dd[1] * t->e1[0] };
    float b2 = (d[0] * s2[0] + d[1] * s2[1] + d[2] * s2[2]) * inv;
    if (b2 < 0 || b1 + b2 >= 1) return 0;
    float tt = (t->e2[0] * s2[0] + t->e2[1] * s2[1] + t->e2[2] * s2[2]) * inv;
    if (tt <= 1e-4f) return 0;
    *tout = tt; return 1;
}
// ordered closest / any hit; counts one "step" per interior node visited (= one child-pair fetch)
static int trace(const float *o, const float *d, float tmax, int any, float *tout)
{
    float id[3] = { 1.f / d[0], 1.f / d[1], 1.f / d[2] };
    int stack[128]; float sent[128]; int sp = 0, hit = -1;
    float e;
    c_nodes++;
    if (!box(&nodes[0], o, id, tmax, &e)) return -1;
    stack[sp] = 0; sent[sp++] = e;
    while (sp) {
        int ni = stack[--sp];
        if (sent[sp] > tmax) continue;
        const Node *n = &nodes[ni];
        if (n->count) {
            for (int k = 0; k < n->count; k++) {
                float t; c_tris++;
                if (tri_hit(&tris[idx[n->first + k]], o, d, &t) && t < tmax) { tmax = t; hit = idx[n->first + k]; if (any) { *tout = t; return hit; } }
            }
            continue;
        }
        c_steps++; c_nodes += 2;
        float el, er;
        int hl = box(&nodes[n->left], o, id, tmax, &el), hr = box(&nodes[n->right], o, id, tmax, &er);
        if (hl && hr) {
            if (er < el) { stack[sp] = n->left; sent[sp++] = el; stack[sp] = n->right; sent[sp++] = er; }
            else { stack[sp] = n->right; sent[sp++] = er; stack[sp] = n->left; sent[sp++] = el; }
        } else if (hl) { stack[sp] = n->left; sent[sp++] = el; }
        else if (hr) { stack[sp] = n->right; sent[sp++] = er; }
    }
    *tout = tmax;
    return hit;
}

static uint64_t rs = 88172645463325252ull;
static float rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (float)((rs >> 40) * (1.0 / 16777216.0)); }

static int depth_of(int n) { if (nodes[n].count) return 0; int a = depth_of(nodes[n].left), b = depth_of(nodes[n].right); return 1 + (a > b ? a : b); }

int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    ntri = (int)(sz / 72);
    double *raw = malloc(sz); if (fread(raw, 1, sz, f) != (size_t)sz) return 1; fclose(f);
    tris = malloc(sizeof(Tri) * ntri); tlo = malloc(12 * ntri); thi = malloc(12 * ntri); tc = malloc(12 * ntri);
    for (int i = 0; i < ntri; i++) for (int a = 0; a < 3; a++) {
        double v0 = raw[9 * i + a], v1 = raw[9 * i + 3 + a], v2 = raw[9 * i + 6 + a];
        tris[i].v0[a] = (float)v0; tris[i].e1[a] = (float)(v1 - v0); tris[i].e2[a] = (float)(v2 - v0);
        tlo[i][a] = (float)fmin(v0, fmin(v1, v2)); thi[i][a] = (float)fmax(v0, fmax(v1, v2));
        tc[i][a] = 0.5f * (tlo[i][a] + thi[i][a]);
    }
    const int W = 384, H = 216;
    const float pos[3] = { 0.f, 0.1f, -2.6f }, tl[3] = { 0.28867513f, 0.26237976f, -2.1f }, rt[3] = { -0.57735027f, 0, 0 }, dn[3] = { 0, -0.32475953f, 0 };
    for (int mode = 0; mode < 4; mode++) {
        int maxleaf = mode == 0 ? 3 : (mode == 1 ? 2 : (mode == 2 ? 3 : 4));
        nodes = malloc(sizeof(Node) * 2 * ntri); nnodes = 0; idx = malloc(sizeof(int) * ntri);
        for (int i = 0; i < ntri; i++) idx[i] = i;
        build(0, ntri, mode ? 1 : 0, maxleaf);
        rs = 88172645463325252ull;
        long long rays[3] = { 0, 0, 0 }, cn[3] = { 0, 0, 0 }, ct[3] = { 0, 0, 0 }, cs[3] = { 0, 0, 0 };
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            float u = (x + rnd()) / W, v = (y + rnd()) / H, o[3], d[3], l = 0;
            for (int a = 0; a < 3; a++) { o[a] = pos[a]; d[a] = tl[a] + u * rt[a] + v * dn[a] - pos[a]; l += d[a] * d[a]; }
            l = 1.f / sqrtf(l); for (int a = 0; a < 3; a++) d[a] *= l;
            for (int b = 0; b < 6; b++) {
                float t; int kind = b ? 1 : 0;
                c_nodes = c_tris = c_steps = 0;
                int h = trace(o, d, 1e8f, 0, &t);
                rays[kind]++; cn[kind] += c_nodes; ct[kind] += c_tris; cs[kind] += c_steps;
                if (h < 0) break;
                const Tri *tr = &tris[h];
                float nrm[3] = { tr->e1[1] * tr->e2[2] - tr->e1[2] * tr->e2[1], tr->e1[2] * tr->e2[0] - tr->e1[0] * tr->e2[2], tr->e1[0] * tr->e2[1] - tr->e1[1] * tr->e2[0] };
                float p[3]; for (int a = 0; a < 3; a++) p[a] = o[a] + t * d[a];
                // shadow ray to a light point
                float lp[3] = { rnd() - 0.5f, 2.5f, rnd() - 0.5f }, sd[3], dist = 0;
                for (int a = 0; a < 3; a++) { sd[a] = lp[a] - p[a]; dist += sd[a] * sd[a]; }
                dist = sqrtf(dist); for (int a = 0; a < 3; a++) sd[a] /= dist;
                c_nodes = c_tris = c_steps = 0;
                float ts; trace(p, sd, dist - 1e-4f, 1, &ts);
                rays[2]++; cn[2] += c_nodes; ct[2] += c_tris; cs[2] += c_steps;
                // uniform hemisphere about the unflipped normal (Material.fs:9-14)
                float w[3];
                for (;;) { w[0] = 2 * rnd() - 1; w[1] = 2 * rnd() - 1; w[2] = 2 * rnd() - 1; float q = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
                    if (q < 1 && q > 1e-6f && w[0] * nrm[0] + w[1] * nrm[1] + w[2] * nrm[2] > 0) { q = 1.f / sqrtf(q); w[0] *= q; w[1] *= q; w[2] *= q; break; } }
                for (int a = 0; a < 3; a++) { o[a] = p[a]; d[a] = w[a]; }
            }
        }
        long long R = rays[0] + rays[1] + rays[2];
        printf("mode %d (%s, leaf<=%d): nodes %d depth %d\n", mode, mode ? "SAH" : "reference median", maxleaf, nnodes, depth_of(0));
        const char *nm[3] = { "primary", "bounce", "shadow" };
        for (int k = 0; k < 3; k++)
            printf("   %-8s rays %8lld  steps/ray %6.2f  nodes/ray %6.2f  tris/ray %6.2f\n", nm[k], rays[k], (double)cs[k] / rays[k], (double)cn[k] / rays[k], (double)ct[k] / rays[k]);
        printf("   all      rays %8lld  steps/ray %6.2f  nodes/ray %6.2f  tris/ray %6.2f   B_ray %.0f\n", R, (double)(cs[0] + cs[1] + cs[2]) / R,
               (double)(cn[0] + cn[1] + cn[2]) / R, (double)(ct[0] + ct[1] + ct[2]) / R,
               32.0 * (cn[0] + cn[1] + cn[2]) / R + 48.0 * (ct[0] + ct[1] + ct[2]) / R + 64);
        free(nodes); free(idx);
    }
    return 0;
}
