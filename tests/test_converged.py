"""Tier T6 on CONVERGED images: the scenes of the five BASELINE configs at 40x40, 256 spp, the config's MAX_DEPTH.

tests/golden/converged_C*.npz hold two renders of each by the REFERENCE itself (oracle/_ref/gi_ref `radiance`: RayTracer::radiance per
pixel through the row loop of run(), two time() seeds A and B; written by `python tests/golden/make_golden.py converged`).  The
implementation under test X (same scene, same sample count, its own counter PRNG and its own photons) must satisfy
    |mean(X) - mean(A, B)| <= 1 % of the mean luminance                      (the bar of SURVEY A.14)
    RMSE(X, A) <= 1.5 * RMSE(A, B) and RMSE(X, B) <= 1.5 * RMSE(A, B)        (X is as close to the reference as the reference to itself)
with radiance clamped to [0, 4] per channel so that single caustic fireflies do not decide the outcome; both numbers are printed."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN, ROOT
from gi_raytracer_b200.abi import render_params

SCENES = {"C1": "cornell", "C2": "caustics", "C3": "glass", "C4": "foliage", "C5": "sponza"}


def _fixture(cfg):
    p = os.path.join(GOLDEN, f"converged_{cfg}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated")
    g = np.load(p)
    res, spp, depth, photons = [int(v) for v in g["meta_res_spp_depth_photons"]]
    return g["radiance_a"], g["radiance_b"], res, spp, depth, photons


def _scene(cfg):
    from gi_raytracer_b200 import host
    sc = host.load_scene(os.path.join(ROOT, "scenes", SCENES[cfg], SCENES[cfg] + ".scn"))
    if sc.n_prims == 0:
        pytest.skip("scene assets not present")
    return sc


def _clamp(a):
    return np.clip(np.nan_to_num(a, nan=0.0, posinf=4.0), 0.0, 4.0)


def _check(x, a, b, what):
    x, a, b = _clamp(x), _clamp(a), _clamp(b)
    rmse = lambda u, v: float(np.sqrt(((u - v) ** 2).mean()))  # noqa: E731
    r_ab, r_xa, r_xb = rmse(a, b), rmse(x, a), rmse(x, b)
    m_ref, m_x = 0.5 * (float(a.mean()) + float(b.mean())), float(x.mean())
    rel = abs(m_x - m_ref) / m_ref
    print(f"{what}: mean luminance ref {m_ref:.6f} X {m_x:.6f} (rel. error {100 * rel:.3f} %, A vs B {100 * abs(float(a.mean()) - float(b.mean())) / m_ref:.3f} %); "
          f"RMSE(A, B) {r_ab:.5f} RMSE(X, A) {r_xa:.5f} RMSE(X, B) {r_xb:.5f}")
    assert rel <= 0.01, (what, m_x, m_ref)
    assert r_xa <= 1.5 * r_ab and r_xb <= 1.5 * r_ab, (what, r_xa, r_xb, r_ab)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4", "C5"])
def test_gpu_converged_image_vs_reference(ctx, cfg):
    a, b, res, spp, depth, photons = _fixture(cfg)
    sc = _scene(cfg)
    ctx.upload_scene(sc)
    if photons:
        ctx.photon_trace(photons, 5, seed=5)
    else:
        ctx.photon_upload(np.zeros((0, 9)))
    ctx.photon_map_build(None)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    _check(acc / spp, a, b, f"GPU, {cfg} {SCENES[cfg]} {res}x{res} {spp} spp depth {depth}")


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_port_converged_image_vs_reference(lib_built, cfg):
    """the CPU restatement under the same protocol (the two cheapest scenes: the CPU suite has to stay short)"""
    a, b, res, spp, depth, photons = _fixture(cfg)
    sc = _scene(cfg)
    ph, _, _ = O.trace_photons(sc, photons, 5, seed=5)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = O.render(sc, O.PMap(ph, sc.root_box), P, 0, 0, res, res, 0, spp)
    _check(acc / spp, a, b, f"oracle port, {cfg}")
