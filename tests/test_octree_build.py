"""Scene octree built on the device (gi_octree_build) against the host build, which is itself pinned
bit for bit to the reference's Octree::rebuild / Node::partition (tests/test_host_scene.py, golden node dumps).
Bar: every array of the flattened tree identical — node boxes (fp64 bits), child index, child mask, leaf ranges, leaf
primitive lists in stored order.  That covers the float SAT of triBoxOverlap (util.cpp:257-330), the sphere / cell test,
cones vanishing from a partitioned tree (entities.h:38-41), the "did not improve" rule and the minimum cell size."""
import os

import numpy as np
import pytest

from conftest import bits_equal, have_assets, scene_path

pytestmark = pytest.mark.gpu

FIELDS = ("node_box", "node_child", "node_mask", "node_prim_off", "node_prim_cnt", "leaf_prims")


def _path(name, synth_dir):
    if name in ("mixed", "cards", "small", "atrium"):
        return os.path.join(synth_dir, name + ".scn")
    if name == "api":
        return os.path.join(synth_dir, "small.scn") + "#api"
    if name in ("cornell", "caustics", "glass") and not have_assets(name):
        pytest.skip(f"assets for {name} not staged")
    p = scene_path(name)
    if name == "foliage" and not os.path.exists(os.path.join(os.path.dirname(p), "cards.obj")):
        pytest.skip("stand-in mesh not generated (scenes/make_standins.py)")
    return p


def _compare(ctx, path):
    from gi_raytracer_b200 import host
    sc = host.load_scene(path)             # host build (the reference's rules)
    boxes = host.prim_boxes(path)
    got, ms = ctx.octree_build(sc.prim_type, sc.prim_geom, boxes, sc.root_box)
    assert got["node_mask"].shape[0] == sc.n_nodes, (got["node_mask"].shape[0], sc.n_nodes)
    assert got["leaf_prims"].shape[0] == sc.leaf_prims.shape[0]
    for f in FIELDS:
        want = getattr(sc, f)
        assert bits_equal(np.ascontiguousarray(got[f]).reshape(want.shape), want), f
    return sc, ms


@pytest.mark.parametrize("name", ["small", "mixed", "cards", "atrium", "api", "cornell", "caustics", "glass", "foliage"])
def test_device_octree_equals_host_octree(ctx, synth_dir, name):
    sc, ms = _compare(ctx, _path(name, synth_dir))
    print(f"{name}: {sc.n_prims} primitives -> {sc.n_nodes} nodes, {sc.leaf_prims.size} leaf references, device build {ms:.2f} ms")


def test_device_octree_large_mesh(ctx):
    """The sponza stand-in: 262 144 triangles -> ~2.0 M nodes / 13.3 M leaf references, 17+ levels."""
    import time
    p = scene_path("sponza")
    if not os.path.exists(os.path.join(os.path.dirname(p), "atrium.obj")):
        pytest.skip("stand-in mesh not generated (scenes/make_standins.py)")
    t0 = time.time()
    sc, ms = _compare(ctx, p)
    print(f"sponza stand-in: {sc.n_prims} triangles -> {sc.n_nodes} nodes, {sc.leaf_prims.size} refs; device build {ms:.1f} ms (host load+build x2 took {time.time() - t0:.1f} s)")
    assert sc.n_nodes > 1000000 and ms < 2000


def test_device_octree_edge_cases(ctx):
    """Root that stays a leaf (<= 16 entities), empty input, entities too thin to be assigned (dx <= 1e-5)."""
    rng = np.random.RandomState(5)
    tri = rng.rand(10, 9)
    typ = np.zeros(10, dtype=np.uint8)
    lo, hi = tri.reshape(10, 3, 3).min(axis=1), tri.reshape(10, 3, 3).max(axis=1) + 3e-5
    got, _ = ctx.octree_build(typ, tri, np.hstack([lo, hi]), np.array([0, 0, 0, 1.1, 1.1, 1.1]))
    assert got["node_mask"].tolist() == [0] and got["node_prim_cnt"].tolist() == [10] and got["leaf_prims"].tolist() == list(range(10))
    got, _ = ctx.octree_build(np.zeros(0, dtype=np.uint8), np.zeros((0, 9)), np.zeros((0, 6)), np.zeros(6))
    assert got["node_mask"].tolist() == [0] and got["node_prim_cnt"].tolist() == [0] and got["leaf_prims"].size == 0
    # 40 triangles lying in planes x = const: their boxes have dx = 3e-5 > 1e-5 and are kept; spheres of radius 1e-6 are dropped
    n = 40
    tri = rng.rand(n, 9)
    tri[:, 0] = tri[:, 3] = tri[:, 6] = np.repeat(rng.rand(n // 2), 2)
    lo, hi = tri.reshape(n, 3, 3).min(axis=1), tri.reshape(n, 3, 3).max(axis=1) + 3e-5
    sph = np.zeros((n, 9)); sph[:, :3] = rng.rand(n, 3); sph[:, 3] = 1e-6
    geom = np.vstack([tri, sph])
    typ = np.concatenate([np.zeros(n, dtype=np.uint8), np.ones(n, dtype=np.uint8)])
    box = np.vstack([np.hstack([lo, hi]), np.hstack([sph[:, :3] - 1e-6, sph[:, :3] + 1e-6])])
    got, _ = ctx.octree_build(typ, geom, box, np.array([0, 0, 0, 1.001, 1.001, 1.001]))
    assert got["node_mask"][0] != 0
    assert (got["leaf_prims"] < n).all() and set(got["leaf_prims"].tolist()) == set(range(n))   # every triangle kept, no sphere
    # leaf ranges tile leaf_prims exactly, interior nodes hold nothing
    leaf = got["node_mask"] == 0
    assert got["node_prim_cnt"][~leaf].sum() == 0 and got["node_prim_cnt"][leaf].sum() == got["leaf_prims"].size
    off = got["node_prim_off"][leaf]; cnt = got["node_prim_cnt"][leaf]
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(cnt)[:-1]]))


def test_cpp_api_renders_identically_with_device_and_host_build(lib_built, synth_dir, monkeypatch, tmp_path):
    """RayTracer::run builds the octree on the device by default; GI_HOST_BUILD=1 keeps the host build: same frame."""
    import ctypes as C
    from gi_raytracer_b200 import capi
    from gi_raytracer_b200.abi import GiStats
    L = capi.load_library()
    path = os.path.join(synth_dir, "mixed.scn").encode()
    imgs = []
    for host_build in (False, True):
        if host_build:
            monkeypatch.setenv("GI_HOST_BUILD", "1")
        else:
            monkeypatch.delenv("GI_HOST_BUILD", raising=False)
        rgb = np.zeros((48 * 48, 3), dtype=np.uint8)
        fs, ps = GiStats(), GiStats()
        pms, fms = C.c_double(), C.c_double()
        rc = L.gih_render_scene(path, 48, 48, 0, 6, 4, 2000, 3, None, rgb.ctypes.data, C.byref(fs), C.byref(ps), C.byref(pms), C.byref(fms))
        assert rc == 0
        imgs.append(rgb)
    assert np.array_equal(imgs[0], imgs[1]) and imgs[0].max() > 0
