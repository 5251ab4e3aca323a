"""Atmosphere: HeightFog::density (atmosphere.h:50-81), Octree::atmosphereDensity / atmosphereBounds
(octree.cpp:214-251), RayTracer::raymarch (raytracer.h:509-529) and the three places it acts (radiance :209-228, visible
:308-316, tracePhotons :658-675).

Tiers: density and bounds are PRNG-free -> bit-exact against the reference's own values (tests/golden/fog_small.npz, written by
oracle/_ref/gi_ref with the noise grid its constructor drew); the march draws one number per step -> GPU vs the oracle with the
shared counter PRNG, bit-exact; radiance with fog -> GPU vs oracle to 1e-6 and both vs the reference statistically, on a scene
whose fog is dense and coloured enough that the same test FAILS for a render without fog."""
import os

import numpy as np
import pytest

import oracle_lib as O
import refdump as R
from conftest import GOLDEN, bits_equal, scene_path, have_assets
from gi_raytracer_b200.abi import SceneArrays, render_params

FOG_SCENE = "caustics_fog_dense"
needs_ref = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/gi_ref not built (make -C oracle ref)")
needs_assets = pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")


@pytest.fixture(scope="module")
def golden_fog():
    return np.load(os.path.join(GOLDEN, "fog_small.npz"))


def _fog_scene(golden_caustics, golden_fog):
    """The caustics geometry of caustics_small.npz with the fog volume and noise grid the reference built."""
    sc = R.scene_from_npz(golden_caustics)
    fp = golden_fog["fog_params_f64"].reshape(-1, 19)
    fogs = np.zeros(fp.shape[0], dtype=SceneArrays.FOG_DTYPE)
    fogs["pos"], fogs["size"], fogs["col"], fogs["density"], fogs["scatter"] = fp[:, 0:3], fp[:, 3:6], fp[:, 6:9], fp[:, 9], fp[:, 10]
    fogs["bmin"], fogs["bmax"] = fp[:, 11:14], fp[:, 14:17]
    fogs["grid_offset"], fogs["grid_count"] = fp[:, 17].astype(np.uint64), fp[:, 18].astype(np.uint64)
    sc.fogs, sc.fog_grid = fogs, np.ascontiguousarray(golden_fog["fog_grid_f64"])
    return sc


def _rays(golden_caustics, golden_fog):
    o, d = golden_caustics["ray_o_f64"].reshape(-1, 3), golden_caustics["ray_d_f64"].reshape(-1, 3)
    bt = golden_fog["fogb_t_f64"].reshape(-1, 3)   # [tmax given, mint out, maxt out]
    return o, d, bt


# ---- CPU: the oracle against the reference's values ---------------------------------------------------------------------------
def test_oracle_fog_density_bit_exact_vs_reference(golden_caustics, golden_fog):
    sc = _fog_scene(golden_caustics, golden_fog)
    dens, col = O.fog_density(sc, golden_fog["fog_pos_f64"].reshape(-1, 3))
    assert (golden_fog["fog_dens_f64"] > 0).mean() > 0.5            # the sample is mostly inside the volume
    assert (golden_fog["fog_dens_f64"] == 0).sum() > 100            # and covers the half-open faces / outside
    assert bits_equal(dens, golden_fog["fog_dens_f64"])
    assert bits_equal(col, golden_fog["fog_col_f64"].reshape(-1, 3))


def test_oracle_atmosphere_bounds_bit_exact_vs_reference(golden_caustics, golden_fog):
    sc = _fog_scene(golden_caustics, golden_fog)
    o, d, bt = _rays(golden_caustics, golden_fog)
    hit, t0, t1 = O.atmosphere_bounds(sc, o, d, bt[:, 0])
    ref_hit = golden_fog["fogb_hit_u8"]
    assert 0.2 < ref_hit.mean() < 0.9
    assert np.array_equal(hit, ref_hit)
    m = ref_hit > 0   # the outputs of a miss are unspecified in the reference (min/max of untouched temporaries)
    assert bits_equal(t0[m], bt[m, 1]) and bits_equal(t1[m], bt[m, 2])
    assert (t0[m] == 0).all()   # octree.cpp:231-244: `min` starts at 0 and can only shrink -> the march starts at the ray origin


def test_host_loader_builds_fog_volume(lib_built):
    """heightFog keyword (sceneLoader.cpp:150-159) -> gi_fog with the reference's box and grid size; the grid is deterministic."""
    from gi_raytracer_b200 import host
    if not have_assets("caustics"):
        pytest.skip("assets not staged")
    a, b = host.load_scene(scene_path(FOG_SCENE)), host.load_scene(scene_path(FOG_SCENE))
    assert a.fogs.size == 1 and np.array_equal(a.fog_grid, b.fog_grid)
    f = a.fogs[0]
    assert np.allclose(f["bmin"], [-4, 0, -4]) and np.allclose(f["bmax"], [4, 4, 4]) and f["density"] == 40
    assert int(f["grid_count"]) == int((8 + 1) * (4 + 1) * (8 + 1) * 2 ** 3)          # atmosphere.h:39
    assert 0 <= a.fog_grid.min() and a.fog_grid.max() < 1 and abs(a.fog_grid.mean() - 0.5) < 0.03


def _fog_check(x, a, b, what):
    import test_reference_radiance as T
    T._check(x, a, b, what)
    # the fog is red: the green/red ratio of the frame separates a fogged render from a clear one
    ratio = lambda u: float(T._clamp(u)[:, 1].mean() / T._clamp(u)[:, 0].mean())  # noqa: E731
    ra, rb, rx = ratio(a), ratio(b), ratio(x)
    print(f"{what}: G/R ratio ref {ra:.3f} {rb:.3f}  X {rx:.3f}")
    assert abs(rx - 0.5 * (ra + rb)) <= max(3 * abs(ra - rb), 0.1), (what, rx, ra, rb)


@needs_ref
@needs_assets
@pytest.mark.ref
def test_oracle_fog_radiance_matches_reference_statistically(lib_built):
    import test_reference_radiance as T
    from gi_raytracer_b200 import host
    res, spp, depth, photons = 48, 32, 8, 100000
    a = T._ref_radiance(FOG_SCENE, res, spp, depth, photons, 1001)
    b = T._ref_radiance(FOG_SCENE, res, spp, depth, photons, 2002)
    sc = host.load_scene(scene_path(FOG_SCENE))
    ph, _, _ = O.trace_photons(sc, photons, 5, seed=5)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = O.render(sc, O.PMap(ph, sc.root_box), P, 0, 0, res, res, 0, spp)
    _fog_check(acc / spp, a, b, "oracle port, fog")
    # the same protocol must reject a render WITHOUT the fog, otherwise it proves nothing
    clear = host.load_scene(scene_path("caustics"))
    ph0, _, _ = O.trace_photons(clear, photons, 5, seed=5)
    acc0, _ = O.render(clear, O.PMap(ph0, clear.root_box), P, 0, 0, res, res, 0, spp)
    with pytest.raises(AssertionError):
        _fog_check(acc0 / spp, a, b, "no fog (must fail)")


# ---- GPU ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_fog_density_and_bounds_bit_exact_vs_reference(ctx, golden_caustics, golden_fog):
    sc = _fog_scene(golden_caustics, golden_fog)
    ctx.upload_scene(sc)
    dens, col = ctx.fog_density(golden_fog["fog_pos_f64"].reshape(-1, 3))
    assert bits_equal(dens, golden_fog["fog_dens_f64"])
    assert bits_equal(col, golden_fog["fog_col_f64"].reshape(-1, 3))
    o, d, bt = _rays(golden_caustics, golden_fog)
    hit, t0, t1, _, _ = ctx.raymarch(o, d, bt[:, 0], march=False)
    assert np.array_equal(hit, golden_fog["fogb_hit_u8"])
    m = hit > 0
    assert bits_equal(t0[m], bt[m, 1]) and bits_equal(t1[m], bt[m, 2])


@pytest.mark.gpu
def test_gpu_raymarch_bit_exact_vs_oracle(ctx, golden_caustics, golden_fog):
    """Same counter PRNG, same sequential position updates: scatter flags, points and colours are identical."""
    sc = _fog_scene(golden_caustics, golden_fog)
    ctx.upload_scene(sc)
    o, d, bt = _rays(golden_caustics, golden_fog)
    tmax = np.where(bt[:, 0] > 1e20, 40.0, bt[:, 0])
    for seed in (1, 77):
        hit, _, _, pos, col = ctx.raymarch(o, d, tmax, seed=seed)
        h2, p2, c2 = O.raymarch(sc, o, d, tmax, seed=seed)
        assert 0.05 < h2.mean() < 0.95
        assert np.array_equal(hit, h2) and bits_equal(pos, p2) and bits_equal(col, c2)
    # squared-distance bound of RayTracer::visible (raytracer.h:308-311) and a segment that ends before the volume
    for tm in (tmax * tmax, np.full_like(tmax, 0.5)):
        hit, _, _, pos, col = ctx.raymarch(o, d, tm, seed=3)
        h2, p2, c2 = O.raymarch(sc, o, d, tm, seed=3)
        assert np.array_equal(hit, h2) and bits_equal(pos, p2)
    # no volumes: never scatters
    clear = R.scene_from_npz(golden_caustics)
    ctx.upload_scene(clear)
    hit, _, _, _, _ = ctx.raymarch(o, d, tmax)
    assert hit.sum() == 0 and ctx.fog_density(o)[0].sum() == 0


@pytest.mark.gpu
@needs_assets
@pytest.mark.parametrize("threshold", ["0", "100000000"], ids=["wavefront", "tail"])
def test_gpu_fog_render_and_photons_vs_oracle(lib_built, monkeypatch, threshold):
    """Fog inside radiance / visible / tracePhotons: GPU (wavefront kernels and tail kernel) vs the oracle, same PRNG."""
    from gi_raytracer_b200 import host
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_TAIL_THRESHOLD", threshold)
    c = Context(0)
    try:
        sc = host.load_scene(scene_path(FOG_SCENE))
        c.upload_scene(sc)
        count = 3000
        n, st = c.photon_trace(count, 5, seed=42)
        got = c.photon_download()
        ref, tries, traces = O.trace_photons(sc, count, 5, seed=42)
        assert n > 0 and abs(int(n) - ref.shape[0]) <= max(3, count // 500)
        assert abs(int(st.photon_tries) - tries) <= max(20, tries // 200)
        if n == ref.shape[0]:
            close = np.isclose(got, ref, rtol=1e-7, atol=1e-9).all(axis=1)
            assert close.mean() > 0.98, close.mean()
        # photons scattered by the fog sit inside the volume, above the floor
        clear = host.load_scene(scene_path("caustics"))
        ref0, _, _ = O.trace_photons(clear, count, 5, seed=42)
        assert (ref[:, 1] > 0.05).mean() > (ref0[:, 1] > 0.05).mean() + 0.05
        c.photon_map_build(None)
        pm = O.PMap(got, sc.root_box)
        res, spp, depth = 40, 4, 8
        P = render_params(res, res, spp, max_depth=depth, seed=99)
        acc, gs = c.render_tile(P, 0, 0, res, res, 0, spp)
        want, os_ = O.render(sc, pm, P, 0, 0, res, res, 0, spp)
        assert abs(int(gs.closest_rays) - int(os_.closest_rays)) <= 0.005 * os_.closest_rays + 4
        assert abs(int(gs.shadow_rays) - int(os_.shadow_rays)) <= 0.005 * os_.shadow_rays + 4
        rel = np.abs(acc - want).max(axis=1) / (np.abs(want).max(axis=1) + 1e-12)
        assert (rel < 1e-6).mean() >= 0.95, (rel < 1e-6).mean()
        assert abs(acc.mean() - want.mean()) <= 0.02 * abs(want.mean())
        # and it is not the clear frame
        c.upload_scene(clear)
        acc0, _ = c.render_tile(P, 0, 0, res, res, 0, spp)
        assert np.abs(acc0 - acc).max(axis=1).mean() > 10 * np.abs(acc - want).max(axis=1).mean() + 1e-6
    finally:
        c.close()


@pytest.mark.gpu
@needs_ref
@needs_assets
def test_gpu_fog_radiance_matches_reference_statistically(ctx):
    import test_reference_radiance as T
    from gi_raytracer_b200 import host
    res, spp, depth, photons = 48, 32, 8, 100000
    a = T._ref_radiance(FOG_SCENE, res, spp, depth, photons, 1001)
    b = T._ref_radiance(FOG_SCENE, res, spp, depth, photons, 2002)
    sc = host.load_scene(scene_path(FOG_SCENE))
    ctx.upload_scene(sc)
    ctx.photon_trace(photons, 5, seed=5)
    ctx.photon_map_build(None)
    P = render_params(res, res, spp, max_depth=depth, seed=11)
    acc, _ = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    _fog_check(acc / spp, a, b, "GPU, fog")
