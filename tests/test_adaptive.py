"""Adaptive sampling (`samples min max thresh`, raytracer.h:100-148): the variance-driven per-pixel sample count.

CPU: the oracle's restatement of the loop against the reference's own `RayTracer::run` image (gi_ref `run`), statistically
(two reference seeds give the yardstick).  GPU: gi_render_adaptive against the oracle (same counter PRNG): identical sample
counts and colours except where an ulp-level libm difference flips a threshold decision."""
import os
import shutil

import numpy as np
import pytest

import oracle_lib as O
import refdump as R
from conftest import ROOT
from gi_raytracer_b200.abi import render_params

SCN = """# adaptive-sampling test scene: cornell subset with samples 4 24
samples 4 24 0.02
photons 60000 5
colorTex 0 0 0
colorTex 1 1 1
colorTex 0.784 0.353 0.404
colorTex 0.404 0.353 0.784
mat 1 0 1 1 1
mat 2 0 1 1 1
mat 3 0 1 1 1
mat 1 0 .0 1 1
mesh assets/test.obj 0 0 0 0 0 0 0
mesh assets/wall_left.obj 0 0 0 0 0 0 1
mesh assets/wall_right.obj 0 0 0 0 0 0 2
mesh assets/sphere.obj -3.5 0 3.35 0 0 0 3
light 0 5 0 4 4 4 .05
"""
MIN_S, MAX_S, THRESH, RES, DEPTH, PHOTONS = 4, 24, 0.02, 56, 4, 60000


@pytest.fixture(scope="module")
def adaptive_scene(tmp_path_factory):
    if not R.have_assets("cornell"):
        pytest.skip("assets not staged")
    d = tmp_path_factory.mktemp("adaptive")
    os.symlink(os.path.join(ROOT, "scenes", "_assets", "cornell"), os.path.join(d, "assets"))
    p = os.path.join(d, "adaptive.scn")
    with open(p, "w") as f:
        f.write(SCN)
    return p


def _resolve8(color):
    c = np.clip(np.power(np.clip(color, 0, None), 1 / 2.2), 0, 1)
    return (255 * c).astype(np.int32)


@pytest.mark.ref
@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/gi_ref not built")
def test_oracle_adaptive_loop_matches_reference_run(lib_built, adaptive_scene):
    from gi_raytracer_b200 import host
    imgs = []
    for tv in (1111, 2222):
        d, _ = R.run_ref(adaptive_scene, ["run"], threads=os.cpu_count(), time_value=tv, w=RES, h=RES, max_depth=DEPTH, photons=PHOTONS)
        imgs.append(R.load(d, "image.u8").reshape(RES * RES, 3).astype(np.int32))
    sc = host.load_scene(adaptive_scene)
    assert sc.knobs["min_samples"] == MIN_S and sc.knobs["max_samples"] == MAX_S
    ph, _, _ = O.trace_photons(sc, PHOTONS, 5, seed=3)
    P = render_params(RES, RES, 1, max_depth=DEPTH, seed=7)
    col, ns = O.render_adaptive(sc, O.PMap(ph, sc.root_box), P, MIN_S, MAX_S, THRESH, 0, 0, RES, RES)
    assert ns.min() >= MIN_S and ns.max() <= MAX_S and len(np.unique(ns)) > 3, "the sample count has to vary over the image"
    x = _resolve8(col)
    mad = lambda u, v: float(np.abs(u - v).mean())  # noqa: E731
    m_ab, m_xa, m_xb = mad(imgs[0], imgs[1]), mad(x, imgs[0]), mad(x, imgs[1])
    print(f"8-bit mean abs diff: ref A vs B {m_ab:.3f}, oracle vs A {m_xa:.3f}, vs B {m_xb:.3f}; samples/pixel {ns.mean():.2f}")
    assert m_ab > 0 and m_xa <= 1.5 * m_ab and m_xb <= 1.5 * m_ab


@pytest.mark.gpu
def test_gpu_adaptive_matches_oracle(ctx, adaptive_scene):
    from gi_raytracer_b200 import host
    sc = host.load_scene(adaptive_scene)
    ctx.upload_scene(sc)
    ctx.photon_trace(PHOTONS, 5, seed=3)
    ph = ctx.photon_download()
    ctx.photon_map_build(None)
    P = render_params(RES, RES, 1, max_depth=DEPTH, seed=7)
    col, ns, st = ctx.render_adaptive(P, MIN_S, MAX_S, THRESH, 0, 0, RES, RES)
    c2, n2 = O.render_adaptive(sc, O.PMap(ph, sc.root_box), P, MIN_S, MAX_S, THRESH, 0, 0, RES, RES)
    same = ns == n2
    assert same.mean() > 0.97, same.mean()
    rel = np.abs(col - c2).max(axis=1) / (np.abs(c2).max(axis=1) + 1e-12)
    assert (rel[same] < 1e-6).mean() > 0.97
    assert int(st.closest_rays) > 0 and ns.min() >= MIN_S and ns.max() <= MAX_S and len(np.unique(ns)) > 3
    # fixed counts are the special case min == max: the running mean equals sum / n up to rounding
    colf, nsf, _ = ctx.render_adaptive(P, 6, 6, THRESH, 0, 0, RES, RES)
    acc, _ = ctx.render_tile(P, 0, 0, RES, RES, 0, 6)
    assert (nsf == 6).all() and np.allclose(colf, acc / 6, rtol=1e-12, atol=1e-15)
