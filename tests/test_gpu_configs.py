"""BASELINE.json configs[2..4] AT THEIR STATED SIZE through the C ABI (C3 Renault 1920x1080 NewPathTracer, C4 99 857
spheres 3840x2160 NewPathTracer, C5 spot x 1708 = 10.0 M triangles 3840x2160 PathIntegrator).

The oracle cannot render these frames in test time, so the bars are: (1) the full-size pixel-centre primary-hit buffers
and 200 k jittered rays hash to the oracle's committed checksums (tests/golden/big_goldens.npz, minted by
make_goldens_big.py) in BOTH precisions -- bit-exact ids and t; (2) size-independent properties of the frames
(determinism, ray-count bounds, tiles summing to the whole frame bit for bit, fast-vs-exact agreement inside the exact
renderer's own seed-to-seed noise)."""
import hashlib

import numpy as np
import pytest

from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, EXACT_F64, FAST_F32

pytestmark = pytest.mark.gpu


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope="module")
def big_scenes():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Scene(scenes.WORKLOADS[name]())
        return cache[name]
    yield get
    for s in cache.values():
        s.close()


@pytest.mark.parametrize("name", ["c4_spheres", "c5_soup"])
def test_full_size_primary_buffers_hash_to_the_oracle(big_goldens, big_scenes, name):
    s = big_scenes(name)
    assert (s.width, s.height) == (3840, 2160)
    uv = np.random.default_rng(1).random((200000, 2))
    for prec in (FAST_F32, EXACT_F64):          # FAST_F32: the hybrid id-exact kernel of the throughput path
        prim, t = s.TracePrimary(precision=prec)
        assert int((prim >= 0).sum()) == int(big_goldens[f"primary/{name}/hits"])
        assert np.array_equal(prim[::997], big_goldens[f"primary/{name}/prim_stride997"])
        assert np.array_equal(t[::997], big_goldens[f"primary/{name}/t_stride997"])
        assert np.array_equal(_sha(prim), big_goldens[f"primary/{name}/sha_prim"])
        assert np.array_equal(_sha(t), big_goldens[f"primary/{name}/sha_t"])
        jp, jt = s.TracePrimary(uv, precision=prec)
        assert int((jp >= 0).sum()) == int(big_goldens[f"jitter/{name}/hits"])
        assert np.array_equal(_sha(jp), big_goldens[f"jitter/{name}/sha_prim"])
        assert np.array_equal(_sha(jt), big_goldens[f"jitter/{name}/sha_t"])


@pytest.mark.parametrize("name,spp", [("c4_spheres", 2), ("c5_soup", 2)])
def test_full_size_frames_properties(big_scenes, name, spp):
    s = big_scenes(name)
    desc = s.desc
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=1)
    a = integ.SampleF32(spp)
    st = dict(integ.stats)
    assert st["paths"] == 3840 * 2160 * spp
    assert st["paths"] <= st["closest_rays"] <= st["paths"] * (desc.max_depth + 1)
    assert st["shadow_rays"] <= st["closest_rays"] and st["hybrid_fixups"] <= 16
    assert np.isfinite(a).all() and a[:, :, :3].mean() > 0
    assert np.array_equal(a, integ.SampleF32(spp))                 # deterministic
    acc = np.zeros_like(a)
    for r in range(4):                                             # 4-way interleaved tiles sum to the frame, bit for bit
        acc += CudaPixelIntegrator(s, precision=FAST_F32, seed=1, tile_size=16, rank=r, world=4).SampleF32(spp)
    assert np.array_equal(acc, a)
    lo = integ.SampleF32(1, first_sample=0).astype(np.float64)
    hi = integ.SampleF32(1, first_sample=1).astype(np.float64)
    assert np.allclose((lo + hi) / 2, a, rtol=1e-5, atol=1e-6)     # progressive frames are exact slices of the sample stream


def test_c3_full_size_mode_b_fast_sits_inside_the_exact_noise():
    """C3 at 1920x1080 under its own integrator (NewPathTracer, mixed Lambert / Metal / SpecularTransmission): specular
    chains diverge path by path between f32 and f64, so the stated tolerance is statistical -- clipped RMSE of fast vs
    exact no larger than exact-vs-exact across seeds, mean radiance within 2e-2."""
    s = Scene(scenes.c3_renault())
    spp = 4
    ea = CudaPixelIntegrator(s, precision=EXACT_F64, seed=5).Sample(spp).copy()[:, :, :3]
    eb = CudaPixelIntegrator(s, precision=EXACT_F64, seed=6).Sample(spp).copy()[:, :, :3]
    fi = CudaPixelIntegrator(s, precision=FAST_F32, seed=5)
    fa = fi.Sample(spp).copy()[:, :, :3]
    assert fi.stats["paths"] == 1920 * 1080 * spp
    clip = np.percentile(ea, 99.0)
    c = lambda x: np.clip(x, -clip, clip)
    noise = np.sqrt(((c(ea) - c(eb)) ** 2).mean())
    err = np.sqrt(((c(fa) - c(ea)) ** 2).mean())
    assert err <= noise, f"fast-vs-exact {err:.3e} exceeds the seed-to-seed noise {noise:.3e}"
    assert abs(c(fa).mean() / c(ea).mean() - 1) < 2e-2
    s.close()
