"""Tier T6: per-pixel radiance against the REFERENCE ITSELF (oracle/_ref/gi_ref = the unmodified sources).

The reference draws from a thread-local, time-seeded xorshift (util.h:52-80); our path uses a counter PRNG.  Radiance can
therefore only agree statistically.  Protocol: render the same small frame with the reference twice (two time seeds, A and
B) and once with the implementation under test X; require
    RMSE(X, A) <= 1.5 * RMSE(A, B)            (X is as close to the reference as the reference is to itself)
    |mean(X) - mean(A)| <= 3 * the spread the two reference runs show, and <= 2 % of the mean luminance
Radiance is clamped to [0, 4] per channel before comparing so that single caustic fireflies do not decide the outcome.
The CPU test pins the oracle port (which the GPU is compared with bit-for-bit elsewhere); the GPU test closes the loop."""
import os

import numpy as np
import pytest

import oracle_lib as O
import refdump as R
from conftest import scene_path
from gi_raytracer_b200.abi import render_params

CASES = [  # scene, resolution, spp, max depth, photons
    ("cornell", 48, 32, 4, 100000),
    ("caustics", 48, 32, 8, 200000),
    ("glass", 40, 32, 16, 50000),
]
needs_ref = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/gi_ref not built (make -C oracle ref)")


def _ref_radiance(name, res, spp, depth, photons, time_value):
    d, meta = R.run_ref(scene_path(name), ["radiance"], threads=os.cpu_count(), time_value=time_value, w=res, h=res, samples=spp, max_depth=depth, photons=photons)
    assert meta["radiance_spp"] == spp
    return R.load(d, "radiance.f64").reshape(res * res, 3)


def _clamp(a):
    return np.clip(np.nan_to_num(a, nan=0.0, posinf=4.0), 0.0, 4.0)


def _check(x, a, b, what):
    x, a, b = _clamp(x), _clamp(a), _clamp(b)
    rmse = lambda u, v: float(np.sqrt(((u - v) ** 2).mean()))  # noqa: E731
    r_ab, r_xa, r_xb = rmse(a, b), rmse(x, a), rmse(x, b)
    m_a, m_b, m_x = float(a.mean()), float(b.mean()), float(x.mean())
    print(f"{what}: RMSE(ref A, ref B) {r_ab:.5f}  RMSE(X, A) {r_xa:.5f}  RMSE(X, B) {r_xb:.5f}  mean A {m_a:.5f} B {m_b:.5f} X {m_x:.5f}")
    assert r_ab > 0, "two reference runs with different seeds cannot be identical"
    assert r_xa <= 1.5 * r_ab and r_xb <= 1.5 * r_ab, (what, r_xa, r_xb, r_ab)
    ref_mean = 0.5 * (m_a + m_b)
    assert abs(m_x - ref_mean) <= max(3 * abs(m_a - m_b), 0.02 * ref_mean), (what, m_x, m_a, m_b)


@needs_ref
@pytest.mark.ref
@pytest.mark.parametrize("name,res,spp,depth,photons", CASES)
def test_oracle_radiance_matches_reference_statistically(lib_built, name, res, spp, depth, photons):
    if not R.have_assets(name):
        pytest.skip("assets not staged")
    from gi_raytracer_b200 import host
    a = _ref_radiance(name, res, spp, depth, photons, 1001)
    b = _ref_radiance(name, res, spp, depth, photons, 2002)
    sc = host.load_scene(scene_path(name))
    ph, _, _ = O.trace_photons(sc, photons, 5, seed=5)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = O.render(sc, O.PMap(ph, sc.root_box), P, 0, 0, res, res, 0, spp)
    _check(acc / spp, a, b, f"oracle port, {name}")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name,res,spp,depth,photons", CASES)
def test_gpu_radiance_matches_reference_statistically(ctx, name, res, spp, depth, photons):
    if not R.have_assets(name):
        pytest.skip("assets not staged")
    from gi_raytracer_b200 import host
    a = _ref_radiance(name, res, spp, depth, photons, 1001)
    b = _ref_radiance(name, res, spp, depth, photons, 2002)
    sc = host.load_scene(scene_path(name))
    ctx.upload_scene(sc)
    ctx.photon_trace(photons, 5, seed=5)
    ctx.photon_map_build(None)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    _check(acc / spp, a, b, f"GPU, {name}")
