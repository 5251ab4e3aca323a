"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE — the CPU restatement used as the checker).
Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from gi_raytracer_b200.abi import GiRenderParams, GiSceneDesc, GiStats, GI_NO_HIT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "gi_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    vp, u64, sz, u32, i32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint32, C.c_int
    L.go_rand.restype = C.c_double
    L.go_rand.argtypes = [u64, u64, u64, u64]
    L.go_halton_sample.restype = C.c_float
    L.go_halton_sample.argtypes = [u32, u32]
    L.go_halton_tables.restype = vp
    L.go_halton_tables.argtypes = [vp]
    L.go_halton_dims.restype = vp
    L.go_henum_init.argtypes = [vp, u32, u32]
    L.go_henum_index.restype = u32
    L.go_henum_index.argtypes = [vp, u32, u32, u32]
    L.go_camera_rays.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    L.go_trace_closest.argtypes = [vp, sz, vp, vp, u64, vp, vp, vp, vp]
    L.go_trace_any.argtypes = [vp, sz, vp, vp, vp, u64, vp]
    L.go_trace_closest_cot.argtypes = [vp, sz, vp, vp, u64, vp, vp, vp, vp]
    L.go_trace_closest_cot_pruned.argtypes = [vp, sz, vp, vp, u64, vp, vp, vp, vp]
    L.go_trace_any_cot.argtypes = [vp, sz, vp, vp, vp, u64, vp, vp, vp]
    L.go_trace_closest_replay.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
    L.go_trace_any_replay.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.go_material_eval.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.go_octree_intersect.argtypes = [vp, sz, vp, vp, vp, vp, u32, vp, vp]
    L.go_octree_intersect_sorted.argtypes = [vp, sz, vp, vp, vp, vp, u32, vp, vp, vp]
    L.go_pmap_build.restype = vp
    L.go_pmap_build.argtypes = [sz, vp, vp]
    L.go_pmap_free.argtypes = [vp]
    L.go_pmap_info.argtypes = [vp, vp, vp, vp, vp]
    L.go_pmap_dump.argtypes = [vp, vp, vp, vp, vp]
    L.go_pmap_candidates.restype = sz
    L.go_pmap_candidates.argtypes = [vp, vp, vp, sz]
    L.go_gather.argtypes = [vp, sz, vp, vp, i32, vp, vp, vp, vp]
    L.go_hemisphere_cos.argtypes = [vp, C.c_float, C.c_float, C.c_double, vp]
    L.go_sphere_cap_cos.argtypes = [vp, C.c_float, C.c_float, C.c_double, C.c_double, vp]
    L.go_sample_phong.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, vp]
    L.go_random_unit_vec.argtypes = [C.c_double, C.c_double, vp]
    L.go_refr.argtypes = [vp, vp, C.c_double, vp]
    L.go_fast_precise_pow.restype = C.c_double
    L.go_fast_precise_pow.argtypes = [C.c_double, C.c_double]
    L.go_trace_photons.restype = sz
    L.go_trace_photons.argtypes = [vp, i32, i32, u64, vp, vp, vp]
    L.go_render.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    L.go_render_adaptive.argtypes = [vp, vp, vp, i32, i32, C.c_double, i32, i32, i32, i32, vp, vp]
    L.go_resolve.argtypes = [sz, vp, i32, vp]
    L.go_child_boxes.argtypes = [vp, vp]
    L.go_fog_density.argtypes = [vp, sz, vp, vp, vp]
    L.go_atmosphere_bounds.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp]
    L.go_raymarch.argtypes = [vp, sz, vp, vp, vp, u64, vp, vp, vp]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data if a is not None else None


def _f64(a, cols):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1, cols)


class HEnum(C.Structure):
    _fields_ = [("p2", C.c_uint32), ("p3", C.c_uint32), ("mx", C.c_uint32), ("my", C.c_uint32), ("inc", C.c_uint32),
                ("w", C.c_uint32), ("h", C.c_uint32), ("scale_x", C.c_float), ("scale_y", C.c_float)]


def halton_sample(dims, idx):
    L = lib()
    dims = np.asarray(dims, dtype=np.uint32).ravel()
    idx = np.asarray(idx, dtype=np.uint32).ravel()
    out = np.empty(dims.size, dtype=np.float32)
    for i in range(dims.size):
        out[i] = L.go_halton_sample(int(dims[i]), int(idx[i]))
    return out


def halton_table():
    """(tables u16[n], dims structured) — the image uploaded to the device is built by the product itself; this is
    only for cross-checking it."""
    L = lib()
    n = C.c_size_t()
    p = L.go_halton_tables(C.byref(n))
    tab = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), shape=(n.value,)).copy()
    dt = np.dtype([("base", "<u4"), ("block", "<u4"), ("nblocks", "<u4"), ("table_off", "<u4"), ("scale", "<f4")])
    d = np.ctypeslib.as_array(C.cast(L.go_halton_dims(), C.POINTER(C.c_uint8)), shape=(256 * dt.itemsize,)).copy().view(dt)
    return tab, d


def henum_index(w, h, s, x, y):
    L = lib()
    he = HEnum()
    L.go_henum_init(C.byref(he), w, h)
    s, x, y = (np.asarray(a, dtype=np.uint32).ravel() for a in (s, x, y))
    return np.array([L.go_henum_index(C.byref(he), int(a), int(b), int(c)) for a, b, c in zip(s, x, y)], dtype=np.uint32)


def camera_rays(scene, w, h, x0, y0, x1, y1, s0, s1):
    L = lib()
    n = (x1 - x0) * (y1 - y0) * (s1 - s0)
    org, d, idx = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, dtype=np.uint32)
    desc = scene.desc()
    L.go_camera_rays(C.byref(desc.camera), w, h, x0, y0, x1, y1, s0, s1, _p(org), _p(d), _p(idx))
    return org, d, idx


def trace_closest(scene, org, d, alpha_seed=0):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    n = org.shape[0]
    prim = np.empty(n, dtype=np.uint32)
    hit, nrm, uv = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 2))
    desc = scene.desc()
    L.go_trace_closest(C.byref(desc), n, _p(org), _p(d), alpha_seed, _p(prim), _p(hit), _p(nrm), _p(uv))
    return prim, hit, nrm, uv


def trace_closest_cot(scene, org, d, alpha_seed=0, pruned=False):
    """canonical ordered traversal with test counts; pruned=True counts what the device executes (same hits)"""
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    n = org.shape[0]
    prim, nn, npr = (np.empty(n, dtype=np.uint32) for _ in range(3))
    hit = np.empty((n, 3))
    desc = scene.desc()
    (L.go_trace_closest_cot_pruned if pruned else L.go_trace_closest_cot)(C.byref(desc), n, _p(org), _p(d), alpha_seed, _p(prim), _p(hit), _p(nn), _p(npr))
    return prim, hit, nn, npr


def trace_any(scene, org, d, maxt2, alpha_seed=0):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    maxt2 = np.ascontiguousarray(maxt2, dtype=np.float64)
    n = org.shape[0]
    vis = np.empty(n, dtype=np.uint8)
    desc = scene.desc()
    L.go_trace_any(C.byref(desc), n, _p(org), _p(d), _p(maxt2), alpha_seed, _p(vis))
    return vis


def trace_any_cot(scene, org, d, maxt2, alpha_seed=0):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    maxt2 = np.ascontiguousarray(maxt2, dtype=np.float64)
    n = org.shape[0]
    vis = np.empty(n, dtype=np.uint8)
    nn, npr = (np.empty(n, dtype=np.uint32) for _ in range(2))
    desc = scene.desc()
    L.go_trace_any_cot(C.byref(desc), n, _p(org), _p(d), _p(maxt2), alpha_seed, _p(vis), _p(nn), _p(npr))
    return vis, nn, npr


def octree_intersect(scene, org, d, tmin, tmax, cap=256):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    n = org.shape[0]
    tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, dtype=np.float64), (n,)))
    tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
    ids = np.full((n, cap), 0xFFFFFFFF, dtype=np.uint32)
    cnt = np.zeros(n, dtype=np.uint32)
    desc = scene.desc()
    L.go_octree_intersect(C.byref(desc), n, _p(org), _p(d), _p(tmin), _p(tmax), cap, _p(ids), _p(cnt))
    return ids, cnt


def octree_intersect_sorted(scene, org, d, tmin, tmax, cap=64):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    n = org.shape[0]
    tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, dtype=np.float64), (n,)))
    tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
    nodes = np.full((n, cap), 0xFFFFFFFF, dtype=np.uint32)
    t0 = np.zeros((n, cap))
    cnt = np.zeros(n, dtype=np.uint32)
    desc = scene.desc()
    L.go_octree_intersect_sorted(C.byref(desc), n, _p(org), _p(d), _p(tmin), _p(tmax), cap, _p(nodes), _p(t0), _p(cnt))
    return nodes, t0, cnt


def trace_closest_replay(scene, org, d, state):
    """RayTracer::trace on the reference's own xorshift64* stream (sequential; `state` = the interposed time() value or the
    state a previous replay call returned) -> (prim, hit, normal, uv, new state)."""
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    n = org.shape[0]
    prim = np.empty(n, dtype=np.uint32)
    hit, nrm, uv = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 2))
    st = C.c_uint64(state)
    desc = scene.desc()
    L.go_trace_closest_replay(C.byref(desc), n, _p(org), _p(d), C.byref(st), _p(prim), _p(hit), _p(nrm), _p(uv))
    return prim, hit, nrm, uv, st.value


def trace_any_replay(scene, org, d, maxt2, state):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    maxt2 = np.ascontiguousarray(maxt2, dtype=np.float64)
    n = org.shape[0]
    vis = np.empty(n, dtype=np.uint8)
    st = C.c_uint64(state)
    desc = scene.desc()
    L.go_trace_any_replay(C.byref(desc), n, _p(org), _p(d), _p(maxt2), C.byref(st), _p(vis))
    return vis, st.value


def material_eval(scene, prim, uv):
    """Material::diffuse->get(uv), emissive->get(uv), Material::getAlpha(uv) of the primitives' materials (material.h)."""
    L = lib()
    prim = np.ascontiguousarray(prim, dtype=np.uint32)
    uv = _f64(uv, 2)
    n = prim.shape[0]
    dif, em, alpha = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
    desc = scene.desc()
    L.go_material_eval(C.byref(desc), n, _p(prim), _p(uv), _p(dif), _p(em), _p(alpha))
    return dif, em, alpha


def fog_density(scene, pos):
    """Octree::atmosphereDensity at points: (density incl. the step-size factor, colour of the last containing volume)."""
    L = lib()
    pos = _f64(pos, 3)
    n = pos.shape[0]
    dens, col = np.empty(n), np.empty((n, 3))
    desc = scene.desc()
    L.go_fog_density(C.byref(desc), n, _p(pos), _p(dens), _p(col))
    return dens, col


def atmosphere_bounds(scene, org, d, tmax):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    tmax = np.ascontiguousarray(tmax, dtype=np.float64)
    n = org.shape[0]
    hit, t0, t1 = np.empty(n, dtype=np.uint8), np.empty(n), np.empty(n)
    desc = scene.desc()
    L.go_atmosphere_bounds(C.byref(desc), n, _p(org), _p(d), _p(tmax), _p(hit), _p(t0), _p(t1))
    return hit, t0, t1


def raymarch(scene, org, d, tmax, seed=1):
    L = lib()
    org, d = _f64(org, 3), _f64(d, 3)
    tmax = np.ascontiguousarray(tmax, dtype=np.float64)
    n = org.shape[0]
    hit, pos, col = np.empty(n, dtype=np.uint8), np.empty((n, 3)), np.empty((n, 3))
    desc = scene.desc()
    L.go_raymarch(C.byref(desc), n, _p(org), _p(d), _p(tmax), seed, _p(hit), _p(pos), _p(col))
    return hit, pos, col


class PMap:
    def __init__(self, photons9, box6):
        self.L = lib()
        self.photons = _f64(photons9, 9)
        self.box = np.ascontiguousarray(box6, dtype=np.float64)
        self.h = self.L.go_pmap_build(self.photons.shape[0], _p(self.photons), _p(self.box))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.go_pmap_free(self.h)
            self.h = None

    def info(self):
        v = [C.c_uint32() for _ in range(4)]
        self.L.go_pmap_info(self.h, *[C.byref(x) for x in v])
        return dict(n_nodes=v[0].value, n_leaves=v[1].value, n_kept=v[2].value, max_depth=v[3].value)

    def dump(self):
        inf = self.info()
        box = np.empty((inf["n_nodes"], 6))
        leaf = np.empty(inf["n_nodes"], dtype=np.uint8)
        cnt = np.empty(inf["n_nodes"], dtype=np.uint32)
        ids = np.empty(inf["n_kept"], dtype=np.uint32)
        self.L.go_pmap_dump(self.h, _p(box), _p(leaf), _p(cnt), _p(ids))
        return box, leaf, cnt, ids

    def candidates(self, pos, cap=65536):
        out = np.empty(cap, dtype=np.uint32)
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = self.L.go_pmap_candidates(self.h, _p(pos), _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def gather(self, pos, d, k=32):
        pos, d = _f64(pos, 3), _f64(d, 3)
        n = pos.shape[0]
        rgb = np.empty((n, 3))
        knn = np.empty((n, k), dtype=np.uint32)
        nc, dl = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
        self.L.go_gather(self.h, n, _p(pos), _p(d), k, _p(rgb), _p(knn), _p(nc), _p(dl))
        return rgb, knn, nc, dl


def trace_photons(scene, count, max_depth=5, seed=1):
    L = lib()
    nl = scene.lights.shape[0]
    buf = np.zeros((max(count * nl, 1), 9))
    tries, traces = C.c_uint64(), C.c_uint64()
    desc = scene.desc()
    n = L.go_trace_photons(C.byref(desc), count, max_depth, seed, _p(buf), C.byref(tries), C.byref(traces))
    return buf[:n].copy(), tries.value, traces.value


def render(scene, pmap, params: GiRenderParams, x0, y0, x1, y1, s0, s1):
    L = lib()
    acc = np.zeros(((y1 - y0) * (x1 - x0), 3))
    st = GiStats()
    desc = scene.desc()
    L.go_render(C.byref(desc), pmap.h if pmap is not None else None, C.byref(params), x0, y0, x1, y1, s0, s1, _p(acc), C.byref(st))
    return acc, st


def render_adaptive(scene, pmap, params: GiRenderParams, min_samples, max_samples, noise_thresh, x0, y0, x1, y1):
    L = lib()
    npx = (y1 - y0) * (x1 - x0)
    col = np.zeros((npx, 3))
    ns = np.zeros(npx, dtype=np.uint32)
    desc = scene.desc()
    L.go_render_adaptive(C.byref(desc), pmap.h if pmap is not None else None, C.byref(params), min_samples, max_samples, float(noise_thresh), x0, y0, x1, y1, _p(col), _p(ns))
    return col, ns


def resolve(accum, spp):
    L = lib()
    accum = _f64(accum, 3)
    out = np.empty((accum.shape[0], 3), dtype=np.uint8)
    L.go_resolve(accum.shape[0], _p(accum), spp, _p(out))
    return out
