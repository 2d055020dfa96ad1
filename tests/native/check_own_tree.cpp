// CPU-only check of the fast path's own tree builder (csrc/mfx_build.cpp); built and run by tests/test_host_abi.py.
// usage: check_own_tree <n> <mode>   mode 0: random small boxes, 1: all boxes identical, 2: a few huge + many small
#include "mfx_build.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>

int mfx_fail(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); return code; }
long mfx_env_long(const char *name, long dflt) { const char *v = getenv(name); return (v && *v) ? atol(v) : dflt; }

static int as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 1000, mode = argc > 2 ? atoi(argv[2]) : 0, max_leaf = 4;
    std::mt19937 rng(n * 7 + mode);
    std::uniform_real_distribution<float> U(-10.f, 10.f), S(0.01f, 0.3f);
    std::vector<float> lo(3 * (size_t)n), hi(3 * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            float c = mode == 1 ? 1.f : U(rng), e = mode == 1 ? 0.5f : S(rng);
            if (mode == 2 && i < 4) { c = 0.f; e = 12.f; }
            lo[3 * (size_t)i + a] = c - e; hi[3 * (size_t)i + a] = c + e;
        }
    MfxOwnTree t;
    mfx_build_own_tree(lo.data(), hi.data(), n, max_leaf, 1.0f, 3, t);
    REQUIRE((int)t.order.size() == n);
    std::vector<int> seen(n, 0);
    for (int v : t.order) { REQUIRE(v >= 0 && v < n); seen[v]++; }
    for (int i = 0; i < n; i++) REQUIRE(seen[i] == 1);
    REQUIRE(!t.quads.empty());
    size_t visited = 0; int depth = 0; long leaf_slots = 0;
    // returns the true bounds of the subtree of record r and checks every stored child box against them
    std::function<bool(int, int, float *, float *)> walk = [&](int r, int level, float *blo, float *bhi) -> bool {
        if (r < 0 || r >= (int)t.quads.size()) return false;
        visited++; depth = std::max(depth, level + 1);
        const QuadF &q = t.quads[r];
        const float *L[3] = { &q.lox.x, &q.loy.x, &q.loz.x }, *H[3] = { &q.hix.x, &q.hiy.x, &q.hiz.x };
        int kids = 0;
        for (int s = 0; s < 4; s++) {
            const int link = as_int((&q.meta.x)[s]);
            if (link == MFX_QUAD_EMPTY) continue;
            kids++;
            float clo[3] = { 3e38f, 3e38f, 3e38f }, chi[3] = { -3e38f, -3e38f, -3e38f };
            if (link >= 0) {
                const int first = link >> 3, cnt = link & 7;
                if (cnt < 1 || cnt > max_leaf || first < 0 || first + cnt > n) return false;
                leaf_slots += cnt;
                for (int k = 0; k < cnt; k++) for (int a = 0; a < 3; a++) {
                    clo[a] = std::min(clo[a], lo[3 * (size_t)t.order[first + k] + a]); chi[a] = std::max(chi[a], hi[3 * (size_t)t.order[first + k] + a]);
                }
            } else if (!walk(~link, level + 1, clo, chi)) return false;
            for (int a = 0; a < 3; a++) {
                if (L[a][s] > clo[a] || H[a][s] < chi[a]) return false;            // stored box must contain its subtree
                blo[a] = std::min(blo[a], clo[a]); bhi[a] = std::max(bhi[a], chi[a]);
            }
        }
        return kids >= 1;
    };
    float blo[3] = { 3e38f, 3e38f, 3e38f }, bhi[3] = { -3e38f, -3e38f, -3e38f };
    REQUIRE(walk(0, 0, blo, bhi));
    REQUIRE(visited == t.quads.size());
    REQUIRE(leaf_slots == n);
    REQUIRE(depth == t.depth);
    printf("ok n=%d mode=%d records=%zu depth=%d\n", n, mode, t.quads.size(), t.depth);
    return 0;
}
