"""GPU parity: the CUDA path (through the C ABI, gi_raytracer_b200.capi) against (1) fixtures produced by the reference
itself and (2) the CPU restatement on the same seeded inputs.  Bars: bit-exact for Halton values/indices, camera rays,
hit primitive ids, hit points / normals / uvs, shadow bits, photon-map cells and k-nearest index sets; 1e-12 relative for the
radiance estimate (summation order is fixed, ties may reorder); stated tolerances for PRNG/libm-dependent radiance."""
import os

import numpy as np
import pytest

import oracle_lib as O
import refdump as R
from conftest import bits_equal, have_assets, scene_path
from gi_raytracer_b200.abi import render_params

pytestmark = pytest.mark.gpu


def random_rays(scene, n, seed):
    """Rays with origins inside the (slightly grown) scene box and uniform directions, normalised like Ray::setDir."""
    rng = np.random.RandomState(seed)
    lo, hi = scene.root_box[:3], scene.root_box[3:]
    pad = 0.1 * (hi - lo)
    o = lo - pad + rng.rand(n, 3) * (hi - lo + 2 * pad)
    d = rng.randn(n, 3)
    d = d * (1.0 / np.sqrt(d[:, 0:1] * d[:, 0:1] + d[:, 1:2] * d[:, 1:2] + d[:, 2:3] * d[:, 2:3]))
    # a few axis-parallel directions: invDir = +-inf and NaN slabs
    d[:8] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0, 1, 0], [1, 0, 0]], dtype=np.float64)
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


# ---- against the reference's own outputs (golden fixtures) ----------------------------------------------------------------
def test_halton_sample_bit_exact(ctx, golden_cornell):
    g = golden_cornell
    idx, val = g["halton_idx_u32"], g["halton_val_f32"]
    dims = np.repeat(np.arange(256, dtype=np.uint32), idx.size)
    got = ctx.halton_sample(dims, np.tile(idx, 256)).reshape(256, -1)
    assert bits_equal(got, val)


def test_halton_index_bit_exact_incl_u32_wrap(ctx, golden_cornell):
    g = golden_cornell
    q, ref = g["henum_query_u32"].reshape(-1, 5), g["henum_index_u32"]
    for w, h in sorted(set((int(a), int(b)) for a, b in q[:, :2])):
        m = (q[:, 0] == w) & (q[:, 1] == h)
        got = ctx.halton_index(w, h, q[m, 2], q[m, 3], q[m, 4])
        assert bits_equal(got, ref[m]), (w, h)


@pytest.mark.parametrize("which", ["cornell", "caustics"])
def test_golden_rays_hits_shadows_gather(ctx, which, golden_cornell, golden_caustics):
    g = golden_cornell if which == "cornell" else golden_caustics
    sc = R.scene_from_npz(g)
    ctx.upload_scene(sc)
    w, h, s0, s1 = [int(v) for v in g["meta_w_h_s0_s1"]]
    o, d, ix = ctx.camera_rays(w, h, 0, 0, w, h, s0, s1)
    ro, rd = g["ray_o_f64"].reshape(-1, 3), g["ray_d_f64"].reshape(-1, 3)
    assert bits_equal(ix, g["ray_idx_u32"]) and bits_equal(o, ro) and bits_equal(d, rd)
    prim, hit, nrm, uv = ctx.trace_closest(ro, rd)
    assert bits_equal(prim, g["hit_id_u32"])
    assert bits_equal(hit, g["hit_pos_f64"].reshape(-1, 3)) and bits_equal(nrm, g["hit_nrm_f64"].reshape(-1, 3)) and bits_equal(uv, g["hit_uv_f64"].reshape(-1, 2))
    vis = ctx.trace_any(g["sh_o_f64"].reshape(-1, 3), g["sh_d_f64"].reshape(-1, 3), g["sh_maxt2_f64"])
    assert bits_equal(vis, g["sh_vis_u8"])
    # photon map built on the device from the reference's photons: identical cells, identical per-leaf photon order
    ctx.photon_upload(g["photons_f64"].reshape(-1, 9))
    ctx.photon_map_build(sc.root_box)
    box, leaf, cnt, ids = ctx.photon_map_download()
    assert bits_equal(box, g["pm_box_f64"].reshape(-1, 6)) and bits_equal(leaf, g["pm_leaf_u8"]) and bits_equal(cnt, g["pm_cnt_u32"]) and bits_equal(ids, g["pm_refs_u32"])
    rgb, knn, nc = ctx.gather(g["q_pos_f64"].reshape(-1, 3), g["q_dir_f64"].reshape(-1, 3), 32)
    assert bits_equal(nc, np.diff(g["q_cand_off_u32"]).astype(np.uint32))
    rknn = g["q_knn_u32"].reshape(-1, 32)
    assert all(set(a) == set(b) for a, b in zip(knn, rknn))
    assert np.allclose(rgb, g["q_est_f64"].reshape(-1, 3), rtol=1e-12, atol=0)


def _golden(name):
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))


@pytest.mark.parametrize("name,scn,kinds_hit", [("api_small", "small.scn#api", {0, 1}), ("cones_small", "small.scn#api2", {0, 1, 2}), ("mixed_small", "mixed.scn", {0, 1})])
def test_golden_sphere_cone_checker_hits_vs_reference(ctx, synth_dir, name, scn, kinds_hit):
    """sphere::intersect / cone::intersect / checkerboard::get (entities.h:60-101, 158-258; material.h:32-49) against fixtures the
    REFERENCE wrote (`gi_ref --api-scene 1|2`, and the `mixed` scene whose semi-opaque materials all have IOR != 1, so its ids are
    PRNG-free): primitive ids and shadow bits bit-exact; hit points / normals / uvs bit-exact on triangles and within 1e-12 on the
    analytic primitives (CUDA's sqrt is IEEE, its atan2 / asin differ from glibc's in the last ulps); texture values bit-exact."""
    from gi_raytracer_b200 import host
    g = _golden(name)
    sc = host.load_scene(os.path.join(synth_dir, scn))
    ctx.upload_scene(sc)
    w, h, s0, s1 = [int(v) for v in g["meta_w_h_s0_s1"]]
    o, d, ix = ctx.camera_rays(w, h, 0, 0, w, h, s0, s1)
    ro, rd = g["ray_o_f64"].reshape(-1, 3), g["ray_d_f64"].reshape(-1, 3)
    assert bits_equal(ix, g["ray_idx_u32"]) and bits_equal(o, ro) and bits_equal(d, rd)
    prim, hit, nrm, uv = ctx.trace_closest(ro, rd)
    rid = g["hit_id_u32"]
    assert bits_equal(prim, rid), f"{(prim != rid).sum()} ids differ from the reference"
    m = rid != 0xFFFFFFFF
    kind = np.full(rid.shape, 255, dtype=np.uint8); kind[m] = sc.prim_type[rid[m]]
    assert set(int(k) for k in np.unique(kind[m])) == kinds_hit
    rh, rn, ruv = g["hit_pos_f64"].reshape(-1, 3), g["hit_nrm_f64"].reshape(-1, 3), g["hit_uv_f64"].reshape(-1, 2)
    tri = kind == 0
    assert bits_equal(hit[tri], rh[tri]) and bits_equal(nrm[tri], rn[tri])
    assert bits_equal(hit[~m], rh[~m])
    ana = m & ~tri
    assert np.allclose(hit[ana], rh[ana], rtol=0, atol=1e-12) and np.allclose(nrm[ana], rn[ana], rtol=0, atol=1e-12)
    assert np.array_equal(np.isnan(uv), np.isnan(ruv)) and np.allclose(uv, ruv, rtol=0, atol=1e-13, equal_nan=True)
    vis = ctx.trace_any(g["sh_o_f64"].reshape(-1, 3), g["sh_d_f64"].reshape(-1, 3), g["sh_maxt2_f64"])
    if name == "mixed_small":   # the reference's shadow test drew from its PRNG only where IOR == 1 and alpha < 1: nowhere in this scene
        pass
    assert bits_equal(vis, g["sh_vis_u8"])
    dif, em, al = ctx.material_eval(rid[m], ruv[m])
    assert bits_equal(dif, g["tex_dif_f64"].reshape(-1, 3)[m]) and bits_equal(em, g["tex_em_f64"].reshape(-1, 3)[m]) and bits_equal(al, g["tex_alpha_f64"][m])


def test_golden_alpha_texture_values_vs_reference(ctx, synth_dir):
    """imageTexture::get / getAlpha and Material::getAlpha (material.h:63-81, 90-93) at the reference's own hit uvs of the
    alpha-card scene: alpha bit-exact, colour = pow(c / 255, 2.2) within 4 ulp of glibc's."""
    from gi_raytracer_b200 import host
    g = _golden("cards_small")
    sc = host.load_scene(os.path.join(synth_dir, "cards_op.scn"))
    ctx.upload_scene(sc)
    rid, ruv = g["hit_id_u32"], g["hit_uv_f64"].reshape(-1, 2)
    m = rid != 0xFFFFFFFF
    dif, em, al = ctx.material_eval(rid[m], ruv[m])
    assert bits_equal(al, g["tex_alpha_f64"][m]) and np.unique(al).size >= 3
    ref = g["tex_dif_f64"].reshape(-1, 3)[m]
    assert np.allclose(dif, ref, rtol=1e-15 * 4, atol=0)
    # the geometric part of the stochastic path is PRNG-free: with every draw passing (alpha_seed irrelevant once opacities are 1)
    # the device and the restatement agree bit for bit on the reference's rays
    ro, rd = g["ray_o_f64"].reshape(-1, 3), g["ray_d_f64"].reshape(-1, 3)
    for seed in (0, 424242):
        a, b = ctx.trace_closest(ro, rd, alpha_seed=seed), O.trace_closest(sc, ro, rd, alpha_seed=seed)
        assert all(bits_equal(x, y) for x, y in zip(a, b))


# ---- against the CPU restatement on seeded inputs ------------------------------------------------------------------------------
def _load(name, synth_dir):
    from gi_raytracer_b200 import host
    if name in ("mixed", "cards", "small", "atrium"):
        return host.load_scene(os.path.join(synth_dir, name + ".scn"))
    if name == "api":   # quadMesh / sphereMesh / coneMesh generators, rotated analytic cones, sphere, box: csrc/host/api_scene.inc
        return host.load_scene(os.path.join(synth_dir, "small.scn") + "#api")
    if not have_assets(name):
        pytest.skip(f"assets for {name} not staged")
    return host.load_scene(scene_path(name))


@pytest.mark.parametrize("name,res", [("mixed", 96), ("cards", 96), ("small", 64), ("atrium", 96), ("api", 96), ("cornell", 128), ("glass", 160)])
def test_closest_and_any_hit_vs_oracle(ctx, synth_dir, name, res):
    sc = _load(name, synth_dir)
    ctx.upload_scene(sc)
    o, d, ix = ctx.camera_rays(res, res, 0, 0, res, res, 0, 2)
    o2, d2, _ = O.camera_rays(sc, res, res, 0, 0, res, res, 0, 2)
    assert bits_equal(o, o2) and bits_equal(d, d2)
    ro, rdir = random_rays(sc, 20000, seed=hash(name) % 1000)
    o, d = np.concatenate([o, ro]), np.concatenate([d, rdir])
    for seed in (0, 12345):
        prim, hit, nrm, uv = ctx.trace_closest(o, d, alpha_seed=seed)
        p2, h2, n2, uv2 = O.trace_closest(sc, o, d, alpha_seed=seed)
        assert bits_equal(prim, p2), f"{name}: {(prim != p2).sum()} ids differ"
        if (sc.prim_type == 2).any():   # cones: sqrt / atan2 of the quadric (CUDA vs glibc, ulps)
            assert np.allclose(hit, h2, rtol=0, atol=1e-12) and np.allclose(nrm, n2, rtol=0, atol=1e-12)
        else:
            assert bits_equal(hit, h2) and bits_equal(nrm, n2)
        # uv: bit-exact for triangle scenes; analytic spheres go through asin/atan2 (CUDA vs glibc differ by ulps), and a
        # normal-less triangle inherits the previous candidate's uv, possibly a sphere's
        if (sc.prim_type == 1).any():   # (the reference's sphereMesh generator emits NaN uvs at the poles: same NaNs on both sides)
            assert np.array_equal(np.isnan(uv), np.isnan(uv2)) and np.allclose(uv, uv2, rtol=0, atol=1e-13, equal_nan=True)
        else:
            assert bits_equal(uv, uv2)
    assert (prim != 0xFFFFFFFF).mean() > 0.2
    # shadow segments from the hits toward the first light (or a fixed point)
    m = prim != 0xFFFFFFFF
    target = sc.lights[0, :3] if sc.lights.shape[0] else sc.root_box[3:] + 1.0
    so = hit[m] + 1e-4 * nrm[m]
    sd = target[None, :] - so
    mt = (sd * sd).sum(axis=1)
    sd = sd * (1.0 / np.sqrt(mt))[:, None]
    for seed in (0, 999):
        vis = ctx.trace_any(so, sd, mt, alpha_seed=seed)
        v2 = O.trace_any(sc, so, sd, mt, alpha_seed=seed)
        assert bits_equal(vis, v2), f"{name}: {(vis != v2).sum()} shadow bits differ"


def test_cone_primitive_vs_oracle(ctx):
    """cones exist only through the C++ API (no .scn keyword) and only survive in an un-partitioned root leaf."""
    from gi_raytracer_b200.abi import SceneArrays
    ang = 0.4
    c, s = np.cos(ang), np.sin(ang)
    rot = np.array([[1, 0, 0], [0, c, s], [0, -s, c]], dtype=np.float64)   # column-major 3x3
    geom = np.zeros((2, 9)); nrm = np.zeros((2, 9))
    geom[0, :5] = [0.0, 0.0, 0.0, 0.8, 2.0]; nrm[0] = rot.T.ravel()
    geom[1, :4] = [1.5, 0.2, 0.3, 0.6]
    mats = np.zeros(1, dtype=SceneArrays.MAT_DTYPE); mats["roughness"] = 1; mats["opacity"] = 1; mats["ior"] = 1
    tex = np.zeros(1, dtype=SceneArrays.TEX_DTYPE); tex["a"] = [[0.5, 0.5, 0.5]]
    cam = np.array([5, 2, 0, -1, 0, 0, 0, 1, 0, 0, 0, 1, 16.8, 9.6], dtype=np.float64)
    sc = SceneArrays(node_box=[[-3, -3, -3, 3, 3, 3]], node_child=[0], node_mask=[0], node_prim_off=[0], node_prim_cnt=[2], leaf_prims=[0, 1],
                     prim_type=[2, 1], prim_geom=geom, prim_nrm=nrm, prim_uv=np.zeros((2, 6)), prim_fnorm=np.zeros((2, 3)), prim_mat=[0, 0], mats=mats, tex=tex,
                     tex_pixels=np.zeros(0, np.uint8), lights=np.zeros((0, 11)), camera=cam, ambient=[0, 0, 0])
    ctx.upload_scene(sc)
    o, d = random_rays(sc, 30000, 3)
    prim, hit, nrm_, uv = ctx.trace_closest(o, d)
    p2, h2, n2, uv2 = O.trace_closest(sc, o, d)
    assert bits_equal(prim, p2)
    assert (prim == 0).sum() > 200 and (prim == 1).sum() > 200
    # cone/sphere hits go through atan2/asin (libm vs CUDA): positions must agree to rounding, uv within 1e-12
    assert np.allclose(hit, h2, rtol=0, atol=1e-12) and np.allclose(nrm_, n2, rtol=0, atol=1e-12) and np.allclose(uv, uv2, rtol=0, atol=1e-12)


def test_photon_map_and_gather_vs_oracle_random_cloud(ctx):
    """Photon clouds with clusters, duplicates and points on cell faces; queries inside and outside the root box."""
    rng = np.random.RandomState(5)
    box = np.array([-2.0, -1.0, -3.0, 2.0, 3.0, 1.0])
    n = 60000
    pos = box[:3] + rng.rand(n, 3) * (box[3:] - box[:3])
    pos[:5000] = np.array([0.3, 0.7, -1.1]) + 0.01 * rng.randn(5000, 3)       # dense cluster
    pos[5000:5040] = np.array([1.0, 1.0, -1.0])                                # > 16 coincident photons: deep degenerate split
    pos[5040:5100, 0] = 0.0                                                    # on the root's x mid-plane
    pos[5100:5110] = box[3:]                                                   # on the max corner: dropped by the half-open test
    ph = np.concatenate([pos, rng.randn(n, 3), rng.rand(n, 3)], axis=1)
    ctx.photon_upload(ph)
    ctx.photon_map_build(box)
    pm = O.PMap(ph, box)
    b1, l1, c1, i1 = ctx.photon_map_download()
    b2, l2, c2, i2 = pm.dump()
    assert ctx.photon_map_info()["n_nodes"] == pm.info()["n_nodes"]
    assert bits_equal(b1, b2) and bits_equal(l1, l2) and bits_equal(c1, c2) and bits_equal(i1, i2)
    q = box[:3] - 0.2 + rng.rand(20000, 3) * (box[3:] - box[:3] + 0.4)
    q[:2000] = pos[rng.randint(0, n, 2000)] + 1e-3 * rng.randn(2000, 3)
    qd = rng.randn(20000, 3)
    def knn_d2(ids, qq):
        ok = ids != 0xFFFFFFFF
        dd = np.full(ids.shape, np.inf)
        pp = pos[np.where(ok, ids, 0)] - qq[:, None, :]
        d2 = pp[..., 0] * pp[..., 0] + pp[..., 1] * pp[..., 1] + pp[..., 2] * pp[..., 2]
        dd[ok] = d2[ok]
        return dd

    for k in (32, 7):
        rgb, knn, nc = ctx.gather(q, qd, k)
        r2, k2, n2, _ = pm.gather(q, qd, k)
        assert bits_equal(nc, n2)
        # exact distance ties (the coincident photons) are unordered in the reference (unstable partial_sort): the selected
        # DISTANCES must agree everywhere, the ids wherever there is no tie
        assert np.array_equal(knn_d2(knn, q), knn_d2(k2, q))
        same = (np.sort(knn, axis=1) == np.sort(k2, axis=1)).all(axis=1)
        assert same.mean() > 0.995, same.mean()
        assert bits_equal(knn[same], k2[same])
        assert np.allclose(rgb[same], r2[same], rtol=1e-12, atol=0)
    # every selection path is exercised: everything selected (<= k), bisection select in registers (k < n <= 256), the
    # streaming sort/merge (> 256 candidates, and the coincident-photon ties)
    assert (nc == 0).sum() > 100 and ((nc > 32) & (nc <= 256)).sum() > 1000 and (nc > 256).any()


def test_empty_and_tiny_photon_maps(ctx):
    box = np.array([0.0, 0.0, 0.0, 1.0, 1.0, 1.0])
    q = np.array([[0.5, 0.5, 0.5], [2.0, 2.0, 2.0]])
    ctx.photon_upload(np.zeros((0, 9)))
    ctx.photon_map_build(box)
    rgb, knn, nc = ctx.gather(q, q, 32)
    assert (nc == 0).all() and (rgb == 0).all() and (knn == 0xFFFFFFFF).all()
    ph = np.random.RandomState(1).rand(9, 9)
    ctx.photon_upload(ph)
    ctx.photon_map_build(box)
    rgb, knn, nc = ctx.gather(q, q, 32)
    r2, k2, n2, _ = O.PMap(ph, box).gather(q, q, 32)
    assert bits_equal(nc, n2) and bits_equal(knn, k2) and np.allclose(rgb, r2, rtol=1e-12, atol=0)
    # sparse maps: fewer candidates than k (everything selected), and k < candidates < 32 (select with idle lanes); query
    # counts that are not multiples of the queries-per-warp batch
    rng = np.random.RandomState(2)
    for nph, nq, k in ((60, 1, 32), (60, 33, 5), (150, 1000, 7), (150, 2049, 3), (400, 5000, 32)):
        ph = rng.rand(nph, 9)
        qq = rng.rand(nq, 3) * 1.2 - 0.1
        ctx.photon_upload(ph)
        ctx.photon_map_build(box)
        rgb, knn, nc = ctx.gather(qq, rng.randn(nq, 3), k)
        ctx.photon_map_build(box)
        qd = rng.randn(nq, 3)
        rgb, knn, nc = ctx.gather(qq, qd, k)
        r2, k2, n2, _ = O.PMap(ph, box).gather(qq, qd, k)
        assert bits_equal(nc, n2) and bits_equal(knn, k2) and np.allclose(rgb, r2, rtol=1e-12, atol=0), (nph, nq, k)
        if nph == 150:
            assert ((nc > k) & (nc < 32)).any()


@pytest.mark.parametrize("name", ["mixed", "cornell"])
def test_photon_trace_vs_oracle(ctx, synth_dir, name):
    sc = _load(name, synth_dir)
    ctx.upload_scene(sc)
    count = 3000
    n, st = ctx.photon_trace(count, 5, seed=42)
    got = ctx.photon_download()
    ref, tries, traces = O.trace_photons(sc, count, 5, seed=42)
    assert n == got.shape[0] and n > 0.5 * count
    # same PRNG, same Halton points; only libm (sin/cos/acos/pow) differs by ulps, so nearly every photon coincides
    assert abs(int(n) - ref.shape[0]) <= max(3, count // 500)
    assert abs(int(st.photon_tries) - tries) <= max(20, tries // 200)
    if n == ref.shape[0]:
        close = np.isclose(got, ref, rtol=1e-7, atol=1e-9).all(axis=1)
        assert close.mean() > 0.99, close.mean()
    # the device map over the device photons equals the CPU map over the same photons
    ctx.photon_map_build(None)
    pm = O.PMap(got, sc.root_box)
    b1, l1, c1, i1 = ctx.photon_map_download()
    b2, l2, c2, i2 = pm.dump()
    assert bits_equal(b1, b2) and bits_equal(i1, i2)


@pytest.mark.parametrize("name,res,spp,depth", [("mixed", 40, 4, 8), ("cards", 32, 4, 4), ("cornell", 48, 4, 4), ("small", 24, 2, 3)])
def test_render_radiance_vs_oracle(ctx, synth_dir, name, res, spp, depth):
    """Same counter PRNG on both sides: per-pixel radiance agrees except where an ulp-level libm difference flips a
    discrete decision.  Tolerance: >= 97 % of pixels within 1e-6 relative, mean luminance within 1 %."""
    sc = _load(name, synth_dir)
    ctx.upload_scene(sc)
    nph = 3000 if sc.knobs.get("photons", 0) > 0 else 0
    pm = None
    if nph:
        ctx.photon_trace(nph, 5, seed=7)
        ph = ctx.photon_download()
        ctx.photon_map_build(None)
        pm = O.PMap(ph, sc.root_box)
    else:
        ctx.photon_upload(np.zeros((0, 9)))
        ctx.photon_map_build(None)
    P = render_params(res, res, spp, max_depth=depth, seed=99)
    acc, st = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    ref, st2 = O.render(sc, pm, P, 0, 0, res, res, 0, spp)
    assert st.closest_rays > 0 and abs(int(st.closest_rays) - int(st2.closest_rays)) <= 0.002 * st2.closest_rays + 4
    assert abs(int(st.shadow_rays) - int(st2.shadow_rays)) <= 0.002 * st2.shadow_rays + 4
    rel = np.abs(acc - ref).max(axis=1) / (np.abs(ref).max(axis=1) + 1e-12)
    assert (rel < 1e-6).mean() >= 0.97, (rel < 1e-6).mean()
    assert abs(acc.mean() - ref.mean()) <= 0.01 * abs(ref.mean()) + 1e-12
    # resolve: 8-bit output within one level (pow differs by ulps)
    img = ctx.resolve(acc, spp)
    img2 = O.resolve(acc, spp)
    assert np.abs(img.astype(int) - img2.astype(int)).max() <= 1


def test_render_tiles_and_sample_ranges_compose(ctx, synth_dir):
    """Size-independent properties: a frame equals the union of its tiles bit-for-bit, and sample ranges add up."""
    sc = _load("mixed", synth_dir)
    ctx.upload_scene(sc)
    ctx.photon_trace(2000, 5, seed=3)
    ctx.photon_map_build(None)
    res, spp = 64, 4
    P = render_params(res, res, spp, max_depth=6, seed=5)
    full, _ = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    again, _ = ctx.render_tile(P, 0, 0, res, res, 0, spp)
    assert bits_equal(full, again)
    full = full.reshape(res, res, 3)
    top, _ = ctx.render_tile(P, 0, 0, res, 24, 0, spp)
    rest, _ = ctx.render_tile(P, 8, 24, 56, res, 0, spp)
    assert bits_equal(top.reshape(24, res, 3), full[:24].copy())
    assert bits_equal(rest.reshape(res - 24, 48, 3), full[24:, 8:56].copy())
    a, _ = ctx.render_tile(P, 0, 0, res, res, 0, 1)
    b, _ = ctx.render_tile(P, 0, 0, res, res, 1, spp)
    assert np.allclose((a + b).reshape(res, res, 3), full, rtol=1e-13, atol=1e-300)


@pytest.mark.skipif(not have_assets("cornell"), reason="assets not staged")
def test_c1_full_size_primary_ids_vs_oracle(ctx):
    """BASELINE config 1 at full size: every primary ray of the 512x512 frame (s = 0): ids and hit points bit-exact."""
    from gi_raytracer_b200 import host
    sc = host.load_scene(scene_path("cornell"))
    ctx.upload_scene(sc)
    o, d, ix = ctx.camera_rays(512, 512, 0, 0, 512, 512, 0, 1)
    prim, hit, nrm, uv = ctx.trace_closest(o, d)
    p2, h2, n2, uv2 = O.trace_closest(sc, o, d)
    assert bits_equal(prim, p2) and bits_equal(hit, h2) and bits_equal(nrm, n2) and bits_equal(uv, uv2)


def test_error_codes_and_empty_inputs(lib_built):
    from gi_raytracer_b200.capi import Context, GiError
    c = Context(0)
    try:
        with pytest.raises(GiError) as e:
            c.trace_closest(np.zeros((4, 3)), np.ones((4, 3)))
        assert e.value.code == -4   # GI_ERR_NO_SCENE
        with pytest.raises(GiError) as e:
            c.gather(np.zeros((4, 3)), np.ones((4, 3)))
        assert e.value.code == -5   # GI_ERR_NO_PHOTONS
        sc = R.scene_from_npz(np.load(os.path.join(os.path.dirname(__file__), "golden", "caustics_small.npz")))
        c.upload_scene(sc)
        prim, hit, nrm, uv = c.trace_closest(np.zeros((0, 3)), np.zeros((0, 3)))
        assert prim.size == 0
        with pytest.raises(GiError) as e:
            c.render_tile(render_params(16, 16, 1), 0, 0, 32, 16, 0, 1)   # tile outside the frame
        assert e.value.code == -1
        bad = R.scene_from_npz(np.load(os.path.join(os.path.dirname(__file__), "golden", "caustics_small.npz")))
        bad.leaf_prims[0] = 10 ** 9
        with pytest.raises(GiError) as e:
            c.upload_scene(bad)
        assert e.value.code == -1
    finally:
        c.close()


def test_device_work_tallies_equal_canonical_traversal_counts(ctx, golden_cornell):
    """The node / primitive test counters the roofline uses are those of the canonical ordered traversal with the
    device's pruning rules R1-R3 (gi_device.cuh; oracle/gi_oracle.c trace_one_cot), counted independently on the CPU for the same
    rays: never more than the reference's own walk, and the same hits."""
    g = golden_cornell
    sc = R.scene_from_npz(g)
    ctx.upload_scene(sc)
    ro, rd = g["ray_o_f64"].reshape(-1, 3), g["ray_d_f64"].reshape(-1, 3)
    prim, hit, _, _ = ctx.trace_closest(ro, rd)
    rays, nn, npr, _ = ctx.last_work("trace_closest")
    p2, h2, cn, cp = O.trace_closest_cot(sc, ro, rd, pruned=True)
    _, _, fn, fp = O.trace_closest_cot(sc, ro, rd)
    assert bits_equal(prim, p2) and bits_equal(hit, h2)
    assert rays == ro.shape[0] and nn == int(cn.sum()) and npr == int(cp.sum())
    assert (cn <= fn).all() and (cp <= fp).all()
    so, sd, mt = g["sh_o_f64"].reshape(-1, 3), g["sh_d_f64"].reshape(-1, 3), g["sh_maxt2_f64"]
    ctx.trace_any(so, sd, mt)
    rays, nn, npr, _ = ctx.last_work("trace_any")
    v, an, ap = O.trace_any_cot(sc, so, sd, mt)
    # any-hit visiting order is free, so only un-blocked rays (which walk everything) have order-independent counts
    assert rays == so.shape[0]
    ctx.trace_any(so[v == 1], sd[v == 1], mt[v == 1])
    rays, nn, npr, _ = ctx.last_work("trace_any")
    assert nn == int(an[v == 1].sum()) and npr == int(ap[v == 1].sum())
    ctx.photon_upload(g["photons_f64"].reshape(-1, 9))
    ctx.photon_map_build(sc.root_box)
    qp, qd = g["q_pos_f64"].reshape(-1, 3), g["q_dir_f64"].reshape(-1, 3)
    ctx.gather(qp, qd, 32)
    q, dsum, csum, ssum = ctx.last_work("gather")
    _, _, nc, dl = O.PMap(g["photons_f64"].reshape(-1, 9), sc.root_box).gather(qp, qd, 32)
    assert q == qp.shape[0] and csum == int(nc.sum()) and ssum == int(np.minimum(nc, 32).sum()) and dsum == int(dl.sum())


@pytest.mark.parametrize("name", ["cards", "mixed", "atrium", "glass"])
def test_pruned_walk_tallies_and_hits_on_alpha_and_mesh_scenes(ctx, synth_dir, name):
    """R1-R3 on scenes where they matter: alpha-textured cards (a fractional-alpha rejection in front of the hit must switch the
    pruning off, raytracer.h:455), spheres / cones, the atrium of large triangles, the glass mesh.  The device's counters equal the
    CPU count of the same rules, the hits equal the UNPRUNED walk of the reference's semantics bit for bit, and the pruning removes work."""
    sc = _load(name, synth_dir)
    ctx.upload_scene(sc)
    o, d, _ = ctx.camera_rays(80, 80, 0, 0, 80, 80, 0, 1)
    ro, rdir = random_rays(sc, 12000, seed=23)
    o, d = np.concatenate([o, ro]), np.concatenate([d, rdir])
    for seed in (0, 777):
        prim, hit, nrm, uv = ctx.trace_closest(o, d, alpha_seed=seed)
        rays, nn, npr, _ = ctx.last_work("trace_closest")
        p1, h1, cn, cp = O.trace_closest_cot(sc, o, d, alpha_seed=seed, pruned=True)
        p0, h0, fn, fp = O.trace_closest_cot(sc, o, d, alpha_seed=seed)
        assert bits_equal(p0, p1) and bits_equal(prim, p0)
        if not (sc.prim_type == 2).any():
            assert bits_equal(h0, h1) and bits_equal(hit, h0)
        assert rays == o.shape[0] and nn == int(cn.sum()) and npr == int(cp.sum()), (name, nn, int(cn.sum()), npr, int(cp.sum()))
        assert (cn <= fn).all() and (cp <= fp).all() and int(cp.sum()) < int(fp.sum())


def test_warp_per_ray_kernels_equal_thread_per_ray(lib_built, synth_dir, monkeypatch):
    """The warp-cooperative traversals (one ray per warp) must return exactly what the thread-per-ray ones return."""
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_TRACE_MODE", "1")
    cw = Context(0)
    monkeypatch.setenv("GI_TRACE_MODE", "0")
    ct = Context(0)
    try:
        for name in ("mixed", "cards", "atrium"):
            sc = _load(name, synth_dir)
            cw.upload_scene(sc); ct.upload_scene(sc)
            o, d, _ = ct.camera_rays(64, 64, 0, 0, 64, 64, 0, 1)
            ro, rdir = random_rays(sc, 15000, seed=11)
            o, d = np.concatenate([o, ro]), np.concatenate([d, rdir])
            a = cw.trace_closest(o, d, alpha_seed=5)
            b = ct.trace_closest(o, d, alpha_seed=5)
            c = O.trace_closest(sc, o, d, alpha_seed=5)
            assert all(bits_equal(x, y) for x, y in zip(a, b)), name
            assert bits_equal(a[0], c[0]) and bits_equal(a[1], c[1])
            assert cw.last_work("trace_closest") == ct.last_work("trace_closest")
            m = a[0] != 0xFFFFFFFF
            so = a[1][m] + 1e-4 * a[2][m]
            sd = sc.lights[0, :3][None, :] - so
            mt = (sd * sd).sum(axis=1)
            sd = sd * (1.0 / np.sqrt(mt))[:, None]
            assert bits_equal(cw.trace_any(so, sd, mt, alpha_seed=9), ct.trace_any(so, sd, mt, alpha_seed=9)), name
    finally:
        cw.close(); ct.close()


def test_tail_kernel_equals_wavefront(lib_built, synth_dir, monkeypatch):
    """Paths finished by the tail megakernel must equal, bit for bit, the same paths run through the wavefront kernels."""
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_TAIL_THRESHOLD", "0")
    c0 = Context(0)
    monkeypatch.setenv("GI_TAIL_THRESHOLD", "100000000")
    c1 = Context(0)
    try:
        for name, depth in (("mixed", 12), ("cards", 5)):
            sc = _load(name, synth_dir)
            outs = []
            for c in (c0, c1):
                c.upload_scene(sc)
                c.photon_trace(2500 if name == "mixed" else 0, 5, seed=2)
                c.photon_map_build(None)
                P = render_params(48, 48, 3, max_depth=depth, seed=17)
                outs.append(c.render_tile(P, 0, 0, 48, 48, 0, 3))
            (a, sa), (b, sb) = outs
            assert bits_equal(a, b), f"{name}: {np.abs(a - b).max()}"
            for f in ("closest_rays", "shadow_rays", "gathers", "closest_node_tests", "closest_prim_tests", "gather_candidates", "gather_selected", "gather_leaf_depth"):
                assert getattr(sa, f) == getattr(sb, f), f
            assert sb.shade_ms > 0 and sa.shade_ms == 0
    finally:
        c0.close(); c1.close()


def test_ray_binning_does_not_change_results(lib_built, synth_dir, monkeypatch):
    """Binning the ray queue by origin cell / direction octant between bounces only permutes the processing order: the
    frame, the ray counts and the work tallies must be bit-identical with and without it."""
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_TAIL_THRESHOLD", "0")
    monkeypatch.setenv("GI_BIN_THRESHOLD", "0")
    c0 = Context(0)
    monkeypatch.setenv("GI_BIN_THRESHOLD", "1")
    c1 = Context(0)
    try:
        for name, depth in (("mixed", 10), ("atrium", 5)):
            sc = _load(name, synth_dir)
            outs = []
            for c in (c0, c1):
                c.upload_scene(sc)
                c.photon_trace(2500 if name == "mixed" else 0, 5, seed=2)
                c.photon_map_build(None)
                P = render_params(64, 48, 3, max_depth=depth, seed=23)
                outs.append(c.render_tile(P, 0, 0, 64, 48, 0, 3))
            (a, sa), (b, sb) = outs
            assert bits_equal(a, b), f"{name}: {np.abs(a - b).max()}"
            for f in ("closest_rays", "shadow_rays", "gathers", "closest_node_tests", "closest_prim_tests", "shadow_node_tests", "shadow_prim_tests", "gather_candidates"):
                assert getattr(sa, f) == getattr(sb, f), f
            assert c1.kernel_ms("bin")[1] > 0 and c0.kernel_ms("bin")[1] == 0
    finally:
        c0.close(); c1.close()


def test_persistent_bounce_kernel_equals_launch_per_queue(lib_built, synth_dir, monkeypatch):
    """k_bounce_p (persistent warps, finished lanes refetch rays) must give the frame, ray counts and work tallies of k_bounce
    (one ray per thread) bit for bit: only the order in which rays are processed differs."""
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_BOUNCE_MODE", "1")
    c1 = Context(0)
    monkeypatch.setenv("GI_BOUNCE_MODE", "2")
    c2 = Context(0)
    try:
        for name, depth in (("mixed", 10), ("atrium", 5), ("cards", 4)):
            sc = _load(name, synth_dir)
            outs = []
            for c in (c1, c2):
                c.upload_scene(sc)
                c.photon_trace(2500 if name == "mixed" else 0, 5, seed=2)
                c.photon_map_build(None)
                P = render_params(72, 40, 3, max_depth=depth, seed=29)
                outs.append(c.render_tile(P, 0, 0, 72, 40, 0, 3))
            (a, sa), (b, sb) = outs
            assert bits_equal(a, b), f"{name}: {np.abs(a - b).max()}"
            for f in ("closest_rays", "shadow_rays", "gathers", "closest_node_tests", "closest_prim_tests", "shadow_node_tests", "shadow_prim_tests", "gather_candidates"):
                assert getattr(sa, f) == getattr(sb, f), f
    finally:
        c1.close(); c2.close()


def test_implicit_child_boxes_equal_loaded_boxes(lib_built, synth_dir, monkeypatch):
    """Traversal with child boxes derived from the parent (the default for trees built by Octree::partition) must equal the
    traversal that loads every child box, bit for bit; reference-built trees must qualify for the implicit path."""
    from gi_raytracer_b200.capi import Context
    monkeypatch.setenv("GI_NO_IMPLICIT_BOXES", "1")
    ce = Context(0)
    monkeypatch.delenv("GI_NO_IMPLICIT_BOXES")
    ci = Context(0)
    try:
        for name in ("atrium", "mixed", "cornell_golden"):
            sc = R.scene_from_npz(np.load(os.path.join(os.path.dirname(__file__), "golden", "cornell_small.npz"))) if name == "cornell_golden" else _load(name, synth_dir)
            ce.upload_scene(sc); ci.upload_scene(sc)
            assert ci.scene_info()["implicit_boxes"] == 1 and ce.scene_info()["implicit_boxes"] == 0
            o, d, _ = ci.camera_rays(64, 64, 0, 0, 64, 64, 0, 1)
            ro, rdir = random_rays(sc, 20000, seed=21)
            o, d = np.concatenate([o, ro]), np.concatenate([d, rdir])
            a, b = ci.trace_closest(o, d, alpha_seed=3), ce.trace_closest(o, d, alpha_seed=3)
            assert all(bits_equal(x, y) for x, y in zip(a, b)), name
            assert ci.last_work("trace_closest") == ce.last_work("trace_closest")
            m = a[0] != 0xFFFFFFFF
            so = a[1][m] + 1e-4 * a[2][m]
            sd = sc.lights[0, :3][None, :] - so
            mt = (sd * sd).sum(axis=1)
            sd = sd * (1.0 / np.sqrt(mt))[:, None]
            assert bits_equal(ci.trace_any(so, sd, mt), ce.trace_any(so, sd, mt)), name
        # a hand-made tree whose child boxes do not follow the formula must fall back to loading them
        sc = _load("mixed", synth_dir)
        sc.node_box[1, 3] += 1e-9
        ci.upload_scene(sc)
        assert ci.scene_info()["implicit_boxes"] == 0
    finally:
        ce.close(); ci.close()


@pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")
def test_c2_full_size_frame_properties(ctx):
    """BASELINE config 2 at its full size (1024x1024, 8 spp, MAX_DEPTH 64, 1 M photons) through size-independent properties:
    the frame equals the union of its tiles bit for bit, sample ranges add up, work tallies are additive, a second render
    is identical, and the photon phase stores exactly the requested number of photons for every light."""
    from gi_raytracer_b200 import host
    sc = host.load_scene(scene_path("caustics"))
    ctx.upload_scene(sc)
    n, pst = ctx.photon_trace(1000000, 5, seed=1)
    assert n == 1000000 * sc.lights.shape[0] and pst.photon_tries >= n
    ctx.photon_map_build(None)
    info = ctx.photon_map_info()
    assert info["n_kept"] <= n and info["n_kept"] > 0.99 * n       # photons on a max face / in an ulp gap are dropped, like the reference
    W = H = 1024
    P = render_params(W, H, 8, max_depth=64, seed=1)
    full, st = ctx.render_tile(P, 0, 0, W, H, 0, 8)
    again, st2 = ctx.render_tile(P, 0, 0, W, H, 0, 8)
    assert bits_equal(full, again) and st.closest_rays == st2.closest_rays and st.closest_node_tests == st2.closest_node_tests
    assert np.isfinite(full).all() and full.mean() > 0   # (a caustic term can be negative: col * dot(photon dir, refDir), raytracer.h:558-576)
    img = full.reshape(H, W, 3)
    top, s_top = ctx.render_tile(P, 0, 0, W, 400, 0, 8)
    bot, s_bot = ctx.render_tile(P, 0, 400, W, H, 0, 8)
    assert bits_equal(top.reshape(400, W, 3), img[:400].copy()) and bits_equal(bot.reshape(H - 400, W, 3), img[400:].copy())
    for f in ("closest_rays", "shadow_rays", "gathers", "closest_node_tests", "closest_prim_tests", "shadow_node_tests", "shadow_prim_tests", "gather_candidates"):
        assert getattr(s_top, f) + getattr(s_bot, f) == getattr(st, f), f
    a, _ = ctx.render_tile(P, 0, 0, W, H, 0, 3)
    b, _ = ctx.render_tile(P, 0, 0, W, H, 3, 8)
    assert np.allclose(a + b, full, rtol=1e-13, atol=1e-300)
    # every primary ray is a closest-hit ray; shadow rays = hits x lights; gathers = hits up to depth 10
    assert st.closest_rays >= W * H * 8 and st.shadow_rays <= st.closest_rays and st.gathers <= st.shadow_rays


@pytest.mark.parametrize("name,w,h,spp,photons,rows", [("glass", 1920, 1080, 64, 275000, 400), ("foliage", 1920, 1080, 16, 0, 500), ("sponza", 3840, 2160, 16, 0, 1200)])
def test_large_configs_tiles_compose_at_full_size(ctx, name, w, h, spp, photons, rows):
    """BASELINE configs 3 (glass, 1920x1080x64, deep specular chains), 4 (foliage stand-in: alpha-textured cards, the FULL traversal
    with its stochastic alpha test; 16 of the 256 spp) and 5 (sponza stand-in, 3840x2160, 16 of the 1024 spp — one GPU of eight
    takes 128 of them; the wrapped range s >= 480 is covered by test_full_size.py) at full resolution: many path chunks, the tail kernel and the side streams all in play; the lower part of
    the frame rendered as its own tile must equal the same rows of the whole frame bit for bit, tallies must add up."""
    from gi_raytracer_b200 import host
    if name == "glass" and not have_assets("glass"):
        pytest.skip("assets not staged")
    p = scene_path(name)
    if name == "foliage" and not os.path.exists(os.path.join(os.path.dirname(p), "cards.obj")):
        pytest.skip("stand-in mesh not generated (scenes/make_standins.py)")
    if name == "sponza" and not os.path.exists(os.path.join(os.path.dirname(p), "atrium.obj")):
        pytest.skip("stand-in mesh not generated (scenes/make_standins.py)")
    sc = host.load_scene(p)
    ctx.upload_scene(sc)
    if photons:
        ctx.photon_trace(photons, 5, seed=1)
    else:
        ctx.photon_upload(np.zeros((0, 9)))
    ctx.photon_map_build(None)
    P = render_params(w, h, spp, max_depth=64, seed=1)
    full, st = ctx.render_tile(P, 0, 0, w, h, 0, spp)
    low, s_low = ctx.render_tile(P, 0, rows, w, h, 0, spp)
    assert np.isfinite(full).all() and full.mean() > 0
    assert bits_equal(low.reshape(h - rows, w, 3), full.reshape(h, w, 3)[rows:].copy())
    up, s_up = ctx.render_tile(P, 0, 0, w, rows, 0, spp)
    assert bits_equal(up.reshape(rows, w, 3), full.reshape(h, w, 3)[:rows].copy())
    for f in ("closest_rays", "shadow_rays", "gathers", "closest_node_tests", "closest_prim_tests", "shadow_node_tests", "shadow_prim_tests"):
        assert getattr(s_up, f) + getattr(s_low, f) == getattr(st, f), f


def test_row_plan_parts_compose_to_the_frame(ctx, synth_dir):
    """gi_render_rows (tile split): the interleaved row blocks of every part, put back at their rows, are the one-call
    frame bit for bit — for part counts that do and do not divide the block count, and a frame height that is not a multiple of the
    block size."""
    sc = _load("mixed", synth_dir)
    ctx.upload_scene(sc)
    ctx.photon_trace(3000, 5, seed=2)
    ctx.photon_map_build(None)
    w, h, spp, block = 96, 77, 3, 16
    P = render_params(w, h, spp, max_depth=6, seed=8)
    full, st = ctx.render_tile(P, 0, 0, w, h, 0, spp)
    full = full.reshape(h, w, 3)
    for nparts in (1, 2, 3, 5, 8):
        frame = np.full((h, w, 3), np.nan)
        rays = 0
        for part in range(nparts):
            if part * block >= h:
                with pytest.raises(Exception):
                    ctx.render_rows(P, block, nparts, part, 0, spp)   # an empty part is refused, not rendered as nothing
                continue
            acc, ps = ctx.render_rows(P, block, nparts, part, 0, spp)
            rows = [y for b in range(part, (h + block - 1) // block, nparts) for y in range(b * block, min(h, (b + 1) * block))]
            assert acc.shape[0] == len(rows) * w == ctx.rows_of_part(h, block, nparts, part) * w
            frame[rows] = acc.reshape(len(rows), w, 3)
            rays += ps.closest_rays + ps.shadow_rays
        assert bits_equal(frame, full), nparts
        assert rays == st.closest_rays + st.shadow_rays


def test_scene_api_queries_vs_reference(ctx, golden_caustics):
    """The batch forms behind the preserved C++ members Octree::intersectSorted / Octree::intersect / PhotonMap::getInRange /
    Entity::intersect (gi_octree_intersect_sorted, gi_octree_intersect, gi_photon_in_range, gi_prim_intersect) against what the
    reference's own members returned for the same rays / points: same leaves in the same order with bit-equal entry distances, same
    entity lists, same candidate photons in the same order, same hit points / normals / uvs."""
    g = golden_caustics
    sc = R.scene_from_npz(g)
    ctx.upload_scene(sc)
    ro, rd = g["ray_o_f64"].reshape(-1, 3), g["ray_d_f64"].reshape(-1, 3)
    nodes, t0, cnt = ctx.octree_intersect_sorted(ro, rd, 0.0, np.inf, cap=64)
    off = g["ls_off_u32"]
    assert bits_equal(cnt, np.diff(off).astype(np.uint32))
    rb, rt = g["ls_box_f64"].reshape(-1, 6), g["ls_t0_f64"]
    for i in range(ro.shape[0]):
        k = int(cnt[i])
        assert bits_equal(sc.node_box[nodes[i, :k]], rb[off[i]:off[i + 1]]) and bits_equal(t0[i, :k], rt[off[i]:off[i + 1]])
    so, sd, mt = g["sh_o_f64"].reshape(-1, 3), g["sh_d_f64"].reshape(-1, 3), g["sh_maxt2_f64"]
    ids, c2 = ctx.octree_intersect(so, sd, 0.0, np.sqrt(mt) - 1e-4, cap=512)
    off, rid = g["sc_off_u32"], g["sc_id_u32"]
    assert bits_equal(c2, np.diff(off).astype(np.uint32))
    for i in range(so.shape[0]):
        assert np.array_equal(ids[i, :c2[i]], rid[off[i]:off[i + 1]])
    # a too small cap reports the full count and fills what fits
    ids4, c4 = ctx.octree_intersect(so[:50], sd[:50], 0.0, np.sqrt(mt[:50]) - 1e-4, cap=4)
    assert bits_equal(c4, c2[:50]) and all(np.array_equal(ids4[i, :min(4, c4[i])], ids[i, :min(4, c4[i])]) for i in range(50))
    # PhotonMap::getInRange: candidate photons in the reference's order
    ctx.photon_upload(g["photons_f64"].reshape(-1, 9))
    ctx.photon_map_build(sc.root_box)
    qp = g["q_pos_f64"].reshape(-1, 3)
    pid, pc = ctx.photon_in_range(qp, cap=1024)
    off, cand = g["q_cand_off_u32"], g["q_cand_u32"]
    assert bits_equal(pc, np.diff(off).astype(np.uint32)) and pc.max() <= 1024
    for i in range(qp.shape[0]):
        assert np.array_equal(pid[i, :pc[i]], cand[off[i]:off[i + 1]])
    # Entity::intersect: the hit primitive of every primary ray, tested on its own
    rid = g["hit_id_u32"]
    m = rid != 0xFFFFFFFF
    ok, hit, nrm, uv, wrote = ctx.prim_intersect(rid[m], ro[m], rd[m])
    assert ok.all() and wrote.all()
    assert bits_equal(hit, g["hit_pos_f64"].reshape(-1, 3)[m]) and bits_equal(nrm, g["hit_nrm_f64"].reshape(-1, 3)[m]) and bits_equal(uv, g["hit_uv_f64"].reshape(-1, 2)[m])
    ok2, *_ = ctx.prim_intersect(np.roll(rid[m], 7), ro[m], rd[m])   # mostly other primitives: mostly misses
    assert ok2.mean() < 0.5
