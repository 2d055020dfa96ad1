"""Parity of the sphere sample's integrator (MFX_SKY_TRACER = GetColor, /root/reference/RenderTest/Sample/
RayTracing.fs:367-382) on the GPU, through the C ABI, against the CPU oracle.

Bars: MFX_EXACT_F64 -- sphere ids, t and radiance bit-exact (the kernel walks a tree where ListHit walks the list:
the answer is the same nearest sphere, ties to the smaller list index); MFX_FAST_F32 -- primary ids differ on at most
2e-4 of the rays, t within 1e-4 relative, frames inside the exact renderer's own seed-to-seed noise with the mean
radiance within 1e-2 (specular chains decorrelate single paths between f32 and f64)."""
import numpy as np
import pytest

from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, Film, EXACT_F64, FAST_F32
from mafrixraytracing_b200 import _lib
from mafrixraytracing_b200.scene import RayTraceCamera
from oracle import oracle
from .test_oracle_sky import MIXED, sky_desc

pytestmark = pytest.mark.gpu


def mixed(width=96, height=64, aperture=0.4, **kw):
    rf, pm = scenes.perlin_tables(3)
    cam = RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, width / height, aperture, 3.4)
    return sky_desc(MIXED["centers"], MIXED["radii"], MIXED["specs"], width=width, height=height, cam=cam, tables=(rf, pm), **kw)


@pytest.mark.parametrize("make", [lambda: scenes.random_scene(), lambda: scenes.random_scene(ground="checker", seed=5), mixed])
def test_exact_primary_and_list_hit_bit_exact(make):
    desc = make()
    s, o = Scene(desc), oracle.OracleSkyScene(desc)
    prim, t = s.TracePrimary(precision=EXACT_F64)
    oprim, ot = o.trace_primary()
    assert np.array_equal(prim, oprim) and np.array_equal(t, ot)
    assert (prim >= 0).mean() > 0.5 and len(np.unique(prim)) > 4
    # ListHit(items, Ray(origin, dir), tmin, tmax) for arbitrary rays: origins in the air, on the ground and inside
    # spheres; unnormalised directions (the Ray constructor normalises)
    rng = np.random.default_rng(1)
    n = 20000
    org = np.concatenate([rng.uniform(-6, 6, (n // 2, 3)) * [1, 0.3, 1] + [0, 1.0, 0],
                          desc.prims["v"][rng.integers(0, len(desc.prims), n // 2), :3] + rng.normal(0, 0.05, (n // 2, 3))])
    dirs = rng.normal(0, 1, (n, 3)) * rng.uniform(0.1, 7.0, (n, 1))
    for (tmin, tmax) in [(0.00001, 10000000.), (0.5, 3.0)]:
        p1, _, t1 = s.Hit(org, dirs, tmin, tmax, precision=EXACT_F64)
        p0, t0 = o.hit(org, dirs, tmin, tmax)
        assert np.array_equal(p1, p0) and np.array_equal(t1, t0)


@pytest.mark.parametrize("make,spp", [(lambda: scenes.random_scene(), 2), (lambda: scenes.random_scene(ground="checker", seed=9, aperture=0.2), 2),
                                      (mixed, 4), (lambda: mixed(aperture=0.0, max_depth=2), 3)])
def test_exact_sample_bit_exact_vs_oracle(make, spp):
    desc = make()
    ref, st = oracle.OracleSkyScene(desc).sample(spp, seed=7, stats=True)
    integ = CudaPixelIntegrator(Scene(desc), precision=EXACT_F64, seed=7)
    tex = integ.Sample(spp)
    assert np.array_equal(tex, ref)
    assert integ.stats["closest_rays"] == st["closest_rays"] and integ.stats["shadow_rays"] == 0
    assert integ.stats["paths"] == desc.width * desc.height * spp
    assert 0.1 < ref[:, :, :3].mean() < 1.0


def test_exact_depth_limit_and_progressive_frames():
    kw = dict(width=16, height=16, cam=RayTraceCamera((0, 0, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.0))
    d0 = sky_desc([(0, 0, 0)], [2.0], [("lambert", (.9, .9, .9))], max_depth=0, **kw)
    tex = CudaPixelIntegrator(Scene(d0), precision=EXACT_F64).Sample(2)
    assert np.all(tex[:, :, :3] == 0.0) and np.all(tex[:, :, 3] == 1.0)
    fast = CudaPixelIntegrator(Scene(d0), precision=FAST_F32).Sample(2)
    assert np.all(fast[:, :, :3] == 0.0)
    # samples are keyed on their absolute index: two half frames average to the whole one
    desc = mixed(width=48, height=32)
    integ = CudaPixelIntegrator(Scene(desc), precision=EXACT_F64, seed=3)
    a = integ.Sample(2, first_sample=0).copy()
    b = integ.Sample(2, first_sample=2).copy()
    whole = oracle.OracleSkyScene(desc).sample(4, seed=3)
    assert np.array_equal(oracle.OracleSkyScene(desc).sample(2, seed=3, first_sample=2), b)
    assert np.allclose((a + b) / 2, whole, rtol=1e-14, atol=0)


@pytest.mark.parametrize("make", [lambda: scenes.random_scene(width=640, height=320), lambda: mixed(width=480, height=320)])
def test_fast_primary_is_bit_exact_and_f32_tests_within_tolerance(monkeypatch, make):
    """Default: the id-exact hybrid kernel (own tree, f32 boxes, f64 Sphere.Hit, ListHit's tie rule) -- ids and t equal
    the oracle's.  MFX_F32_PRIMARY=1: the f32 sphere tests of the deeper bounces -- ids within 2e-4, t within 1e-4."""
    desc = make()
    s = Scene(desc)
    o = oracle.OracleSkyScene(desc)
    oprim, ot = o.trace_primary()
    rng = np.random.default_rng(2)
    org = rng.uniform(-6, 6, (20000, 3)) * [1, 0.2, 1] + [0, 1.5, 0]      # arbitrary rays, origins off every surface
    dirs = rng.normal(0, 1, (20000, 3)) * 3.0
    p0, t0 = o.hit(org, dirs)
    prim, t = s.TracePrimary(precision=FAST_F32)
    assert np.array_equal(prim, oprim) and np.array_equal(t, ot)
    p1, _, t1 = s.Hit(org, dirs, 0.00001, 10000000., precision=FAST_F32)
    assert np.array_equal(p1, p0) and np.array_equal(t1[p0 >= 0], t0[p0 >= 0])
    monkeypatch.setenv("MFX_F32_PRIMARY", "1")
    prim, t = s.TracePrimary(precision=FAST_F32)
    differ = prim != oprim
    assert differ.mean() <= 2e-4, f"{differ.sum()} of {differ.size} sphere ids differ"
    both = (~differ) & (oprim >= 0)
    assert np.max(np.abs(t[both] - ot[both]) / ot[both]) <= 1e-4
    p1, _, t1 = s.Hit(org, dirs, 0.00001, 10000000., precision=FAST_F32)
    assert (p1 != p0).mean() <= 2e-4
    ok = (p1 == p0) & (p0 >= 0)
    assert np.max(np.abs(t1[ok] - t0[ok]) / t0[ok]) <= 1e-4


@pytest.mark.parametrize("make,flags", [(lambda: scenes.random_scene(width=200, height=100), _lib.SAMPLE_REFERENCE_STREAM),
                                        (lambda: scenes.random_scene(width=200, height=100), 0),
                                        (lambda: mixed(width=120, height=80), _lib.SAMPLE_REFERENCE_STREAM),
                                        (lambda: mixed(width=120, height=80), 0)])
def test_fast_sample_is_statistically_equal_to_exact(make, flags):
    """Both samplers of the fast path -- the reference's rejection loops on the exact mode's stream, and the default
    direct draws (uniform ball point without the loop) -- against the f64 frame."""
    desc = make()
    s = Scene(desc)
    spp = 64
    ea = CudaPixelIntegrator(s, precision=EXACT_F64, seed=5).Sample(spp).copy()[:, :, :3]
    eb = CudaPixelIntegrator(s, precision=EXACT_F64, seed=6).Sample(spp).copy()[:, :, :3]
    fi = CudaPixelIntegrator(s, precision=FAST_F32, seed=5)
    fa = fi.Sample(spp, flags=flags).copy()[:, :, :3]
    noise = np.sqrt(((ea - eb) ** 2).mean())
    err = np.sqrt(((fa - ea) ** 2).mean())
    assert err <= 1.05 * noise, f"fast-vs-exact {err:.3e} exceeds the seed-to-seed noise {noise:.3e}"
    assert abs(fa.mean() / ea.mean() - 1) < 1e-2
    assert fi.stats["shadow_rays"] == 0 and fi.stats["launches_shadow"] == 0
    assert fa.min() >= 0.0 and fa.max() <= 1.0 + 1e-6           # attenuations <= 1, sky <= 1


def test_fast_matched_stream_paths_take_the_same_decisions():
    """A diffuse-only scene (no specular chains): on the matched stream single paths agree, not just their statistics."""
    centers = [(0, -100.5, -1), (0, 0, -1), (-1.05, 0, -1), (1.05, 0, -1)]
    specs = [("checker", (0.2, 0.3, 0.1), (0.9, 0.9, 0.9)), ("lambert", (0.8, 0.3, 0.3)), ("lambert", (0.2, 0.4, 0.9)), ("lambert", (0.7, 0.7, 0.7))]
    desc = sky_desc(centers, [100.0, 0.5, 0.5, 0.5], specs, width=160, height=96,
                    cam=RayTraceCamera((0.2, 0.6, 2.5), (0, 0, -1), (0, 1, 0), 45.0, 160 / 96))
    s = Scene(desc)
    exact = CudaPixelIntegrator(s, precision=EXACT_F64, seed=5).Sample(16).copy()[:, :, :3]
    fast = CudaPixelIntegrator(s, precision=FAST_F32, seed=5).Sample(16, flags=_lib.SAMPLE_REFERENCE_STREAM).copy()[:, :, :3]
    rel = np.sqrt(((fast - exact) ** 2).mean()) / exact.mean()
    assert rel <= 2e-2, f"relative RMSE {rel:.3e}"
    assert abs(fast.mean() / exact.mean() - 1) < 2e-3


@pytest.mark.parametrize("precision", [EXACT_F64, FAST_F32])
def test_tile_sharding_and_film_in_sky_mode(precision):
    desc = mixed(width=100, height=60)
    s = Scene(desc)
    whole = CudaPixelIntegrator(s, precision=precision, seed=2).Sample(3).copy()
    parts = [CudaPixelIntegrator(s, precision=precision, seed=2, tile_size=16, rank=r, world=3).Sample(3).copy() for r in range(3)]
    assert np.array_equal(sum(parts)[:, :, :3], whole[:, :, :3])
    film = Film(s)
    integ = CudaPixelIntegrator(s, precision=precision, seed=2)
    f1 = film.GetFrame(integ, 2).copy()
    f2 = film.GetFrame(integ, 2).copy()
    assert np.allclose(f2[:, :, :3], (f1[:, :, :3] + integ.Sample(2, first_sample=2)[:, :, :3]) / 2, rtol=1e-6, atol=1e-7)
    film.close()


@pytest.mark.parametrize("name", ["random_scene", "mixed"])
def test_exact_matches_the_committed_sky_goldens_and_display_transform(name):
    """The committed vectors (tests/golden/sky_goldens.npz) and the sphere sample's display transform
    (sqrt, int(255.99 c), vertical flip -- RayTracing.fs:456-460) through Film.PostProcess."""
    import hashlib
    from .test_oracle_sky import _sky_goldens, golden_sky_scenes
    g = _sky_goldens()
    full, small = golden_sky_scenes()[name]
    prim, t = Scene(full).TracePrimary(precision=EXACT_F64)
    assert np.array_equal(prim, g[f"primary/{name}/prim"]) and np.array_equal(t[::97], g[f"primary/{name}/t_stride97"])
    s = Scene(small)
    integ = CudaPixelIntegrator(s, precision=EXACT_F64, seed=7)
    film = Film(s)
    tex = film.GetFrame(integ, 2, first_sample=0).copy()
    assert np.array_equal(tex[:, :, :3], g[f"image/{name}/rgb"])
    disp = film.PostProcess()
    assert np.array_equal(disp, oracle.sky_display_rgba8(tex))
    assert np.array_equal(np.frombuffer(hashlib.sha256(disp.tobytes()).digest(), dtype=np.uint8), g[f"image/{name}/display_sha"])
    film.close()


@pytest.mark.parametrize("make", [lambda: scenes.random_scene(width=320, height=160), lambda: mixed(width=240, height=160)])
def test_late_bounces_in_one_launch_give_the_same_frame(monkeypatch, make):
    """From bounce MFX_SKY_TAIL (default 8) on the fast path traces AND shades every surviving path to the depth limit
    in one launch (k_f_trace6<TAIL>); the wavefront loop (MFX_SKY_TAIL=0: two launches per bounce up to depth 50) must
    give the same frame bit for bit -- both call the one sky_scatter_f -- and count the same rays, with a fraction of
    the launches.  A tail that starts right after the camera rays (MFX_SKY_TAIL=1) must agree too."""
    desc = make()
    s = Scene(desc)
    integ = CudaPixelIntegrator(s, precision=FAST_F32, seed=4)
    monkeypatch.setenv("MFX_SKY_TAIL", "0")
    ref = integ.Sample(4).copy()
    st0 = dict(integ.stats)
    for tail in ("8", "1", "3"):
        monkeypatch.setenv("MFX_SKY_TAIL", tail)
        img = integ.Sample(4).copy()
        st = dict(integ.stats)
        assert np.array_equal(img, ref), tail
        assert st["closest_rays"] == st0["closest_rays"], (tail, st["closest_rays"], st0["closest_rays"])
        assert st["launches"] < st0["launches"] // 2
    assert np.isfinite(ref).all() and ref[..., :3].mean() > 0.05
