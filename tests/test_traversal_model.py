"""Host-side check of the traversal ALGORITHM the fast kernel runs (stackless bit-trail over the
heap-indexed tree, ray replacement, vote-postponed leaves): a lock-step 32-lane CPU model of
k_f_trace4 (tools/sim_warp.py) must finish every ray and agree with the oracle's exhaustive
Bvh.CheckHit on primitive ids -- for any refill / vote thresholds."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import ROOT
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sim_warp  # noqa: E402
from mafrixraytracing_b200 import scenes  # noqa: E402
from oracle import oracle  # noqa: E402


@pytest.mark.parametrize("name,kw", [("c2_spot", dict(width=20, height=12)), ("cornell", dict(width=14, height=14)),
                                     ("c1_cube", dict(width=16, height=12))])
@pytest.mark.parametrize("refill_t,leaf_t", [(1, 1), (12, 16), (1, 32), (32, 4)])
def test_warp_model_finishes_and_matches_oracle(name, kw, refill_t, leaf_t):
    desc = scenes.WORKLOADS[name](**kw)
    sc = sim_warp.flatten(desc)
    uv = []
    rays = sim_warp.camera_rays(desc, seed=3, uv_out=uv)
    out, iters, stuck = sim_warp.run_warp(sc, rays, refill_t, leaf_t, max_iters=20000)
    assert stuck is None and len(out) == len(rays)
    prim, t = oracle.OracleScene(desc).trace_primary(np.array(uv))
    got = np.array([sc["ref"][sc["slots"][out[r][1]][-1]] if out[r][1] >= 0 else -1 for r in range(len(rays))])
    assert (got != prim).mean() <= 0.02        # f32 boxes/rays vs the f64 oracle: only edge-grazing rays may differ
    both = (got == prim) & (prim >= 0)
    tt = np.array([out[r][0] for r in range(len(rays))])
    assert np.allclose(tt[both], t[both], rtol=1e-4)


def test_warp_model_shadow_queries_terminate():
    desc = scenes.c2_spot(width=16, height=10)
    sc = sim_warp.flatten(desc)
    rays = [(o, d, 2.5) for (o, d, _) in sim_warp.camera_rays(desc)]
    out, iters, stuck = sim_warp.run_warp(sc, rays, 12, 16, ANY=True, max_iters=20000)
    assert stuck is None and len(out) == len(rays)
    o = oracle.OracleScene(desc)
    op, _, _ = o.hit([r[0] for r in rays], [r[1] for r in rays], 1e-6, 2.5)
    occl = np.array([out[r][1] >= 0 for r in range(len(rays))])
    assert (occl != (op >= 0)).mean() <= 0.02
