"""compute-sanitizer is closed on the GPU pool, so the library carries its own checks: `make DEBUG=1` builds
libmafrix_cuda_dbg.so with a bounds check on every index the wavefront kernels compute (records, slots, stack entries, queue
positions, path ids, materials, pixels) and the assertion the non-atomic `rad` update of the shadow kernel rests on -- at
most one shadow ray per path and bounce.  A violation makes Sample fail.  This test runs the every-kernel script
(tools/sanitize_run.py: all integrators, both precisions, the hybrid kernel and its fixup, counters, seams, tiles, stripes,
async, multi-GPU) against that build, and then proves the assertion is alive by injecting a fault."""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
DBG = os.path.join(ROOT, "mafrixraytracing_b200", "libmafrix_cuda_dbg.so")


def _run(code, **env):
    e = dict(os.environ, MFX_LIB=DBG, **env)
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_every_kernel_passes_the_debug_checks():
    if not os.path.exists(DBG):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "mafrixraytracing_b200", "csrc"), "-s", "DEBUG=1"])
    r = _run("import runpy; from mafrixraytracing_b200 import _lib; assert b'DEBUG CHECKS' in _lib.load().mfx_version(); "
             "runpy.run_path('tools/sanitize_run.py', run_name='__main__')")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok ") == 6


def test_the_shadow_ray_assertion_fires_on_an_injected_fault():
    code = ("from mafrixraytracing_b200 import scenes, Scene, CudaPixelIntegrator, FAST_F32, MafrixError\n"
            "i = CudaPixelIntegrator(Scene(scenes.cornell(width=64, height=64)), precision=FAST_F32, seed=1)\n"
            "i.Sample(2)\n"                                    # stamps start at zero: clean
            "try:\n    i.Sample(2)\n    print('NOT CAUGHT')\n"  # stamps of the first call left in place: every shadow ray is a 'second' one
            "except MafrixError as e:\n    print('CAUGHT', e)\n")
    r = _run(code, MFX_DEBUG_FAULT="1")
    assert r.returncode == 0 and "CAUGHT" in r.stdout and "debug build" in r.stdout, r.stdout + r.stderr
    r = _run(code)                                             # without the fault both calls pass
    assert "NOT CAUGHT" in r.stdout, r.stdout + r.stderr
