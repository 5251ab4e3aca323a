"""ctypes binding of libgi_b200.so — the C ABI declared in include/gi_api.h.

`Context` wraps one gi_ctx (one CUDA device).  Host-pointer methods take/return numpy arrays; `*_dev` methods take raw
device pointers (e.g. torch tensors' data_ptr()).  There is no CPU fallback: if the library or a B200 is missing the
constructor raises.
"""
import ctypes as C
import os

import numpy as np

from .abi import GiRenderParams, GiSceneDesc, GiStats, SceneArrays

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgi_b200.so")

# every symbol include/gi_api.h declares (tests check that the library exports each one)
API_SYMBOLS = [
    "gi_create", "gi_destroy", "gi_last_error", "gi_version", "gi_stream", "gi_synchronize", "gi_scene_upload",
    "gi_halton_sample", "gi_halton_index", "gi_camera_rays", "gi_trace_closest", "gi_trace_closest_dev", "gi_trace_any",
    "gi_trace_any_dev", "gi_photon_trace", "gi_photon_upload", "gi_photon_count", "gi_photon_download",
    "gi_photon_map_build", "gi_photon_map_info", "gi_photon_map_download", "gi_photon_map_slab_size",
    "gi_photon_map_slab_ptr", "gi_photon_map_adopt_slab", "gi_photon_map_reserve_slab", "gi_photon_gather",
    "gi_photon_gather_dev", "gi_render_tile", "gi_render_tile_dev", "gi_resolve", "gi_resolve_dev", "gi_last_kernel_ms",
    "gi_last_work", "gi_scene_info", "gi_render_adaptive", "gi_render_adaptive_dev", "gi_fog_density", "gi_raymarch", "gi_octree_build", "gi_octree_download", "gi_cancel", "gi_render_image", "gi_configure", "gi_material_eval",
    "gi_rows_of_part", "gi_render_rows", "gi_render_rows_dev", "gi_comm_unique_id", "gi_comm_init", "gi_comm_destroy", "gi_comm_info", "gi_photon_map_bcast",
    "gi_framebuffer_reduce", "gi_framebuffer_gather", "gi_render_rows_image", "gi_octree_intersect", "gi_octree_intersect_sorted", "gi_photon_in_range",
    "gi_prim_intersect",
]

_LIB = None


def comm_unique_id():
    """gi_comm_unique_id: 128 bytes to hand to every rank's Context.comm_init (made on ONE rank)."""
    L = load_library()
    buf = (C.c_uint8 * 128)()
    rc = L.gi_comm_unique_id(buf, 128)
    if rc != 0:
        raise GiError(rc, "gi_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf)


class GiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gi error {code}: {msg}")
        self.code = code


def load_library():
    """dlopen libgi_b200.so (built in-tree by gi_raytracer_b200.build).  Fails loudly when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    lib_path = os.environ.get("GI_LIB", LIB_PATH)   # A/B runs of kernel variants (profiles/frame_ab.py); default: the in-tree build
    if not os.path.exists(lib_path):
        raise FileNotFoundError(f"{lib_path} is missing: run `python -m gi_raytracer_b200.build` (there is no CPU fallback)")
    L = C.CDLL(lib_path)
    vp, u64, sz, u32, i32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint32, C.c_int
    L.gi_create.argtypes = [i32, C.POINTER(vp)]
    L.gi_destroy.argtypes = [vp]
    L.gi_destroy.restype = None
    L.gi_last_error.argtypes = [vp]
    L.gi_last_error.restype = C.c_char_p
    L.gi_version.restype = C.c_char_p
    L.gi_stream.argtypes = [vp]
    L.gi_stream.restype = vp
    L.gi_synchronize.argtypes = [vp]
    L.gi_scene_upload.argtypes = [vp, C.POINTER(GiSceneDesc)]
    L.gi_halton_sample.argtypes = [vp, sz, vp, vp, vp]
    L.gi_halton_index.argtypes = [vp, i32, i32, sz, vp, vp, vp, vp]
    L.gi_camera_rays.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    for f in (L.gi_trace_closest, L.gi_trace_closest_dev):
        f.argtypes = [vp, sz, vp, vp, u64, vp, vp, vp, vp]
    for f in (L.gi_trace_any, L.gi_trace_any_dev):
        f.argtypes = [vp, sz, vp, vp, vp, u64, vp]
    L.gi_fog_density.argtypes = [vp, sz, vp, vp, vp]
    L.gi_material_eval.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    def sig(name, argtypes):
        # GI_LIB names an A/B build of the library (profiles/ab_variants.sh), possibly older than this binding: a symbol it lacks is
        # skipped there; the in-tree library must export every one (tests/test_abi.py)
        try:
            getattr(L, name).argtypes = argtypes
        except AttributeError:
            if "GI_LIB" not in os.environ:
                raise
    sig("gi_rows_of_part", [i32, i32, i32, i32])
    for f in ("gi_render_rows", "gi_render_rows_dev"):
        sig(f, [vp, C.POINTER(GiRenderParams), i32, i32, i32, i32, i32, vp, C.POINTER(GiStats)])
    sig("gi_comm_unique_id", [vp, sz])
    sig("gi_comm_init", [vp, vp, sz, i32, i32])
    sig("gi_comm_destroy", [vp])
    sig("gi_comm_info", [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)])
    sig("gi_photon_map_bcast", [vp, i32])
    sig("gi_framebuffer_reduce", [vp, vp, sz, i32])
    sig("gi_framebuffer_gather", [vp, vp, sz, i32, i32, vp, i32])
    sig("gi_render_rows_image", [vp, C.POINTER(GiRenderParams), i32, i32, i32, vp, i32, C.POINTER(GiStats)])
    L.gi_raymarch.argtypes = [vp, sz, vp, vp, vp, u64, i32, vp, vp, vp, vp, vp]
    sig("gi_octree_intersect", [vp, sz, vp, vp, vp, vp, u32, vp, vp])
    sig("gi_octree_intersect_sorted", [vp, sz, vp, vp, vp, vp, u32, vp, vp, vp])
    sig("gi_photon_in_range", [vp, sz, vp, u32, vp, vp])
    sig("gi_prim_intersect", [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp])
    L.gi_octree_build.argtypes = [vp, u32, vp, vp, vp, vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_double)]
    L.gi_octree_download.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.gih_scene_prim_bbox.argtypes = [vp, vp]
    L.gih_scene_rebuild_device.argtypes = [vp, vp, C.POINTER(C.c_double)]
    L.gi_photon_trace.argtypes = [vp, i32, i32, u64, C.POINTER(u64), C.POINTER(GiStats)]
    L.gi_photon_upload.argtypes = [vp, sz, vp]
    L.gi_photon_count.argtypes = [vp, C.POINTER(sz)]
    L.gi_photon_download.argtypes = [vp, sz, vp]
    L.gi_photon_map_build.argtypes = [vp, vp]
    L.gi_photon_map_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.gi_photon_map_download.argtypes = [vp, vp, vp, vp, vp]
    L.gi_photon_map_slab_size.argtypes = [vp, C.POINTER(sz)]
    L.gi_photon_map_slab_ptr.argtypes = [vp, C.POINTER(vp)]
    L.gi_photon_map_adopt_slab.argtypes = [vp, sz]
    L.gi_photon_map_reserve_slab.argtypes = [vp, sz, C.POINTER(vp)]
    for f in (L.gi_photon_gather, L.gi_photon_gather_dev):
        f.argtypes = [vp, sz, vp, vp, i32, vp, vp, vp]
    for f in (L.gi_render_tile, L.gi_render_tile_dev):
        f.argtypes = [vp, C.POINTER(GiRenderParams), i32, i32, i32, i32, i32, i32, vp, C.POINTER(GiStats)]
    for f in (L.gi_render_adaptive, L.gi_render_adaptive_dev):
        f.argtypes = [vp, C.POINTER(GiRenderParams), i32, i32, C.c_double, i32, i32, i32, i32, vp, vp, C.POINTER(GiStats)]
    for f in (L.gi_resolve, L.gi_resolve_dev):
        f.argtypes = [vp, sz, vp, i32, vp]
    L.gi_last_kernel_ms.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(u64)]
    L.gi_last_work.argtypes = [vp, C.c_char_p, C.POINTER(u64 * 4)]
    L.gi_scene_info.argtypes = [vp, C.POINTER(u32 * 4)]
    # host-side scene facade (same library)
    L.gih_scene_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.gih_scene_desc.argtypes = [vp]
    L.gih_scene_desc.restype = C.POINTER(GiSceneDesc)
    L.gih_scene_knobs.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double)]
    L.gih_scene_knobs.restype = None
    L.gih_scene_free.argtypes = [vp]
    L.gih_scene_free.restype = None
    L.gih_render_scene.argtypes = [C.c_char_p, i32, i32, i32, i32, i32, i32, u64, C.c_char_p, vp, C.POINTER(GiStats), C.POINTER(GiStats),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.gih_render_progressive.argtypes = [C.c_char_p, i32, i32, i32, i32, i32, i32, u64, i32, i32, vp, C.POINTER(i32), C.POINTER(C.c_double)]
    L.gi_cancel.argtypes = [vp, i32]
    L.gi_configure.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.gi_render_image.argtypes = [vp, C.POINTER(GiRenderParams), i32, i32, i32, i32, i32, i32, vp, vp, C.POINTER(GiStats)]
    L.gih_png_decode.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp, sz]
    L.gih_png_encode.argtypes = [C.c_char_p, i32, i32, vp]
    L.gih_jpg_decode.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), vp, sz]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data if a is not None else None


def _f64(a, cols):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1, cols)


class Context:
    """One gi_ctx on one CUDA device."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.gi_create(device, C.byref(h))
        if rc != 0:
            raise GiError(rc, "gi_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device
        self._scene_keepalive = None

    def close(self):
        if getattr(self, "h", None):
            self.L.gi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GiError(rc, self.L.gi_last_error(self.h).decode())

    @property
    def stream(self):
        return self.L.gi_stream(self.h)

    def synchronize(self):
        self._ck(self.L.gi_synchronize(self.h))

    def kernel_ms(self, family):
        ms, n = C.c_double(), C.c_uint64()
        self._ck(self.L.gi_last_kernel_ms(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def scene_info(self):
        out = (C.c_uint32 * 4)()
        self._ck(self.L.gi_scene_info(self.h, C.byref(out)))
        return dict(full=int(out[0]), implicit_boxes=int(out[1]), n_nodes=int(out[2]), n_leaf_refs=int(out[3]))

    def last_work(self, family):
        out = (C.c_uint64 * 4)()
        self._ck(self.L.gi_last_work(self.h, family.encode(), C.byref(out)))
        return [int(v) for v in out]

    # -- scene ------------------------------------------------------------------------------------------------
    def upload_scene(self, scene: SceneArrays):
        d = scene.desc()
        self._ck(self.L.gi_scene_upload(self.h, C.byref(d)))

    # -- Halton / camera ------------------------------------------------------------------------------------------
    def halton_sample(self, dims, idx):
        dims = np.ascontiguousarray(dims, dtype=np.uint32).ravel()
        idx = np.ascontiguousarray(idx, dtype=np.uint32).ravel()
        out = np.empty(dims.size, dtype=np.float32)
        self._ck(self.L.gi_halton_sample(self.h, dims.size, _p(dims), _p(idx), _p(out)))
        return out

    def halton_index(self, w, h, s, x, y):
        s, x, y = (np.ascontiguousarray(a, dtype=np.uint32).ravel() for a in (s, x, y))
        out = np.empty(s.size, dtype=np.uint32)
        self._ck(self.L.gi_halton_index(self.h, w, h, s.size, _p(s), _p(x), _p(y), _p(out)))
        return out

    def camera_rays(self, w, h, x0, y0, x1, y1, s0, s1):
        n = (x1 - x0) * (y1 - y0) * (s1 - s0)
        org, d, idx = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, dtype=np.uint32)
        self._ck(self.L.gi_camera_rays(self.h, w, h, x0, y0, x1, y1, s0, s1, _p(org), _p(d), _p(idx)))
        return org, d, idx

    # -- rays --------------------------------------------------------------------------------------------------------
    def trace_closest(self, org, d, alpha_seed=0):
        org, d = _f64(org, 3), _f64(d, 3)
        n = org.shape[0]
        prim = np.empty(n, dtype=np.uint32)
        hit, nrm, uv = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 2))
        self._ck(self.L.gi_trace_closest(self.h, n, _p(org), _p(d), alpha_seed, _p(prim), _p(hit), _p(nrm), _p(uv)))
        return prim, hit, nrm, uv

    def trace_any(self, org, d, maxt2, alpha_seed=0):
        org, d = _f64(org, 3), _f64(d, 3)
        maxt2 = np.ascontiguousarray(maxt2, dtype=np.float64).ravel()
        n = org.shape[0]
        vis = np.empty(n, dtype=np.uint8)
        self._ck(self.L.gi_trace_any(self.h, n, _p(org), _p(d), _p(maxt2), alpha_seed, _p(vis)))
        return vis

    def material_eval(self, prim, uv):
        """Material::diffuse->get(uv), emissive->get(uv), Material::getAlpha(uv) for the materials of the given primitives."""
        prim = np.ascontiguousarray(prim, dtype=np.uint32).ravel()
        uv = _f64(uv, 2)
        n = prim.size
        dif, em, alpha = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
        self._ck(self.L.gi_material_eval(self.h, n, _p(prim), _p(uv), _p(dif), _p(em), _p(alpha)))
        return dif, em, alpha

    # -- the reference's scene-API queries, batch forms -----------------------------------------------------------------------------
    def octree_intersect(self, org, d, tmin, tmax, cap=256):
        """Octree::intersect -> (prim ids [n, cap] (valid up to min(count, cap)), counts [n])."""
        org, d = _f64(org, 3), _f64(d, 3)
        n = org.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, dtype=np.float64), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
        ids = np.full((n, cap), 0xFFFFFFFF, dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        self._ck(self.L.gi_octree_intersect(self.h, n, _p(org), _p(d), _p(tmin), _p(tmax), cap, _p(ids), _p(cnt)))
        return ids, cnt

    def octree_intersect_sorted(self, org, d, tmin, tmax, cap=64):
        """Octree::intersectSorted -> (flattened node ids [n, cap], entry distances [n, cap], counts [n])."""
        org, d = _f64(org, 3), _f64(d, 3)
        n = org.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, dtype=np.float64), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
        nodes = np.full((n, cap), 0xFFFFFFFF, dtype=np.uint32)
        t0 = np.zeros((n, cap))
        cnt = np.zeros(n, dtype=np.uint32)
        self._ck(self.L.gi_octree_intersect_sorted(self.h, n, _p(org), _p(d), _p(tmin), _p(tmax), cap, _p(nodes), _p(t0), _p(cnt)))
        return nodes, t0, cnt

    def photon_in_range(self, pos, cap=512):
        """PhotonMap::getInRange -> (original photon ids [n, cap], counts [n])."""
        pos = _f64(pos, 3)
        n = pos.shape[0]
        ids = np.full((n, cap), 0xFFFFFFFF, dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        self._ck(self.L.gi_photon_in_range(self.h, n, _p(pos), cap, _p(ids), _p(cnt)))
        return ids, cnt

    def prim_intersect(self, prim, org, d):
        """Entity::intersect of primitive prim[i] with ray i -> (ok, hit, normal, uv, wrote_uv)."""
        prim = np.ascontiguousarray(prim, dtype=np.uint32).ravel()
        org, d = _f64(org, 3), _f64(d, 3)
        n = prim.size
        ok, wrote = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
        hit, nrm, uv = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 2))
        self._ck(self.L.gi_prim_intersect(self.h, n, _p(prim), _p(org), _p(d), _p(ok), _p(hit), _p(nrm), _p(uv), _p(wrote)))
        return ok, hit, nrm, uv, wrote

    def fog_density(self, pos):
        """Octree::atmosphereDensity at points -> (density incl. the step-size factor, colour of the last containing volume)."""
        pos = _f64(pos, 3)
        n = pos.shape[0]
        dens, col = np.empty(n), np.empty((n, 3))
        self._ck(self.L.gi_fog_density(self.h, n, _p(pos), _p(dens), _p(col)))
        return dens, col

    def raymarch(self, org, d, tmax, seed=1, march=True):
        """Octree::atmosphereBounds (mint 0, maxt tmax) and, with march, RayTracer::raymarch -> (hit, t0, t1, pos, col)."""
        org, d = _f64(org, 3), _f64(d, 3)
        tmax = np.ascontiguousarray(tmax, dtype=np.float64)
        n = org.shape[0]
        hit, t0, t1 = np.empty(n, dtype=np.uint8), np.empty(n), np.empty(n)
        pos, col = np.empty((n, 3)), np.empty((n, 3))
        self._ck(self.L.gi_raymarch(self.h, n, _p(org), _p(d), _p(tmax), seed, 1 if march else 0, _p(hit), _p(t0), _p(t1), _p(pos), _p(col)))
        return hit, t0, t1, pos, col

    def cancel(self, raise_=True):
        """gi_cancel: callable from any thread while another thread is inside a render / photon call on this context."""
        self._ck(self.L.gi_cancel(self.h, 1 if raise_ else 0))

    def configure(self, key, value):
        """gi_configure: a scheduling knob (overlap_threshold, sched_mode, ring, tail_threshold, bin_threshold, bounce_mode, trace_mode, tail_mode)."""
        self._ck(self.L.gi_configure(self.h, key.encode(), int(value)))

    def octree_build(self, prim_type, prim_geom, prim_bbox, root_box):
        """Octree::rebuild / Node::partition on the device -> (dict of gi_scene_desc node arrays, device ms)."""
        prim_type = np.ascontiguousarray(prim_type, dtype=np.uint8)
        prim_geom, prim_bbox = _f64(prim_geom, 9), _f64(prim_bbox, 6)
        root_box = np.ascontiguousarray(root_box, dtype=np.float64)
        nn, nr, ms = C.c_uint32(), C.c_uint32(), C.c_double()
        self._ck(self.L.gi_octree_build(self.h, prim_type.size, _p(prim_type), _p(prim_geom), _p(prim_bbox), _p(root_box), C.byref(nn), C.byref(nr), C.byref(ms)))
        out = dict(node_box=np.empty((nn.value, 6)), node_child=np.empty(nn.value, dtype=np.uint32), node_mask=np.empty(nn.value, dtype=np.uint8),
                   node_prim_off=np.empty(nn.value, dtype=np.uint32), node_prim_cnt=np.empty(nn.value, dtype=np.uint32), leaf_prims=np.empty(nr.value, dtype=np.uint32))
        self._ck(self.L.gi_octree_download(self.h, _p(out["node_box"]), _p(out["node_child"]), _p(out["node_mask"]), _p(out["node_prim_off"]), _p(out["node_prim_cnt"]),
                                           _p(out["leaf_prims"]) if nr.value else None))
        return out, ms.value

    def trace_closest_dev(self, n, org_ptr, dir_ptr, prim_ptr, hit_ptr=None, nrm_ptr=None, uv_ptr=None, alpha_seed=0):
        self._ck(self.L.gi_trace_closest_dev(self.h, n, org_ptr, dir_ptr, alpha_seed, prim_ptr, hit_ptr, nrm_ptr, uv_ptr))

    def trace_any_dev(self, n, org_ptr, dir_ptr, maxt2_ptr, vis_ptr, alpha_seed=0):
        self._ck(self.L.gi_trace_any_dev(self.h, n, org_ptr, dir_ptr, maxt2_ptr, alpha_seed, vis_ptr))

    # -- photons -------------------------------------------------------------------------------------------------------
    def photon_trace(self, count, max_depth=5, seed=1):
        n, st = C.c_uint64(), GiStats()
        self._ck(self.L.gi_photon_trace(self.h, count, max_depth, seed, C.byref(n), C.byref(st)))
        return n.value, st

    def photon_upload(self, photons9):
        ph = _f64(photons9, 9)
        self._ck(self.L.gi_photon_upload(self.h, ph.shape[0], _p(ph)))

    def photon_count(self):
        n = C.c_size_t()
        self._ck(self.L.gi_photon_count(self.h, C.byref(n)))
        return n.value

    def photon_download(self):
        n = self.photon_count()
        out = np.empty((n, 9))
        self._ck(self.L.gi_photon_download(self.h, n, _p(out)))
        return out

    def photon_map_build(self, box6=None):
        b = np.ascontiguousarray(box6, dtype=np.float64) if box6 is not None else None
        self._ck(self.L.gi_photon_map_build(self.h, _p(b)))

    def photon_map_info(self):
        v = [C.c_uint32() for _ in range(4)]
        self._ck(self.L.gi_photon_map_info(self.h, *[C.byref(x) for x in v]))
        return dict(n_nodes=v[0].value, n_leaves=v[1].value, n_kept=v[2].value, max_depth=v[3].value)

    def photon_map_download(self):
        inf = self.photon_map_info()
        box = np.empty((inf["n_nodes"], 6))
        leaf = np.empty(inf["n_nodes"], dtype=np.uint8)
        cnt = np.empty(inf["n_nodes"], dtype=np.uint32)
        ids = np.empty(max(inf["n_kept"], 1), dtype=np.uint32)
        self._ck(self.L.gi_photon_map_download(self.h, _p(box), _p(leaf), _p(cnt), _p(ids)))
        return box, leaf, cnt, ids[:inf["n_kept"]]

    def photon_map_slab(self):
        """(device pointer, bytes) of the built map's slab — what rank 0 broadcasts."""
        n, p = C.c_size_t(), C.c_void_p()
        self._ck(self.L.gi_photon_map_slab_size(self.h, C.byref(n)))
        self._ck(self.L.gi_photon_map_slab_ptr(self.h, C.byref(p)))
        return p.value, n.value

    def photon_map_reserve_slab(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.gi_photon_map_reserve_slab(self.h, nbytes, C.byref(p)))
        return p.value

    def photon_map_adopt_slab(self, nbytes):
        self._ck(self.L.gi_photon_map_adopt_slab(self.h, nbytes))

    def gather(self, pos, d, k=32, want_knn=True):
        pos, d = _f64(pos, 3), _f64(d, 3)
        n = pos.shape[0]
        rgb = np.empty((n, 3))
        knn = np.empty((n, k), dtype=np.uint32) if want_knn else None
        nc = np.empty(n, dtype=np.uint32)
        self._ck(self.L.gi_photon_gather(self.h, n, _p(pos), _p(d), k, _p(rgb), _p(knn), _p(nc)))
        return rgb, knn, nc

    def gather_dev(self, n, pos_ptr, dir_ptr, rgb_ptr, k=32, knn_ptr=None, ncand_ptr=None):
        self._ck(self.L.gi_photon_gather_dev(self.h, n, pos_ptr, dir_ptr, k, rgb_ptr, knn_ptr, ncand_ptr))

    # -- frame -----------------------------------------------------------------------------------------------------------
    def render_tile(self, params: GiRenderParams, x0, y0, x1, y1, s0, s1):
        acc = np.empty(((y1 - y0) * (x1 - x0), 3))
        st = GiStats()
        self._ck(self.L.gi_render_tile(self.h, C.byref(params), x0, y0, x1, y1, s0, s1, _p(acc), C.byref(st)))
        return acc, st

    def render_image(self, params: GiRenderParams, x0, y0, x1, y1, s0, s1, want_accum=False):
        """gi_render_image: the resolved 8-bit frame (and optionally the fp64 sums), one call, no accumulator round trip."""
        npx = (x1 - x0) * (y1 - y0)
        rgb = np.empty((npx, 3), dtype=np.uint8)
        acc = np.empty((npx, 3)) if want_accum else None
        st = GiStats()
        self._ck(self.L.gi_render_image(self.h, C.byref(params), x0, y0, x1, y1, s0, s1, _p(rgb), _p(acc) if want_accum else None, C.byref(st)))
        return rgb, acc, st

    def render_adaptive(self, params: GiRenderParams, min_samples, max_samples, noise_thresh, x0, y0, x1, y1):
        """gi_render_adaptive -> (colour [npx, 3] fp64 running means, samples taken [npx] u32, GiStats)."""
        npx = (x1 - x0) * (y1 - y0)
        col = np.empty((npx, 3), dtype=np.float64)
        ns = np.empty(npx, dtype=np.uint32)
        st = GiStats()
        self._ck(self.L.gi_render_adaptive(self.h, C.byref(params), min_samples, max_samples, float(noise_thresh), x0, y0, x1, y1, col.ctypes.data, ns.ctypes.data, C.byref(st)))
        return col, ns, st

    # -- tile split / multi-GPU ---------------------------------------------------------------------------------------------
    def rows_of_part(self, height, block_rows, nparts, part):
        n = self.L.gi_rows_of_part(height, block_rows, nparts, part)
        if n < 0:
            raise GiError(n, "bad row plan")
        return n

    def render_rows(self, params: GiRenderParams, block_rows, nparts, part, s0, s1):
        """gi_render_rows: the rows of one part of the interleaved row-block plan -> (accum [local rows * width, 3], GiStats)."""
        rows = self.rows_of_part(params.height, block_rows, nparts, part)
        acc = np.empty((rows * params.width, 3))
        st = GiStats()
        self._ck(self.L.gi_render_rows(self.h, C.byref(params), block_rows, nparts, part, s0, s1, _p(acc), C.byref(st)))
        return acc, st

    def render_rows_dev(self, params: GiRenderParams, block_rows, nparts, part, s0, s1, accum_ptr):
        st = GiStats()
        self._ck(self.L.gi_render_rows_dev(self.h, C.byref(params), block_rows, nparts, part, s0, s1, accum_ptr, C.byref(st)))
        return st

    def comm_init(self, id_bytes, rank, nranks):
        """gi_comm_init: collective over the nranks contexts; id_bytes = the 128 bytes of comm_unique_id() made on one rank."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id_bytes)[:128].ljust(128, b"\0"))
        self._ck(self.L.gi_comm_init(self.h, buf, 128, rank, nranks))

    def comm_destroy(self):
        self._ck(self.L.gi_comm_destroy(self.h))

    def comm_info(self):
        r, n, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.gi_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v)))
        return dict(rank=r.value, nranks=n.value, nccl_version=v.value)

    def photon_map_bcast(self, root=0):
        self._ck(self.L.gi_photon_map_bcast(self.h, root))

    def framebuffer_reduce(self, accum_ptr, count, root=0):
        self._ck(self.L.gi_framebuffer_reduce(self.h, accum_ptr, count, root))

    def framebuffer_gather(self, local_ptr, row_bytes, height, block_rows, frame_ptr, root=0):
        self._ck(self.L.gi_framebuffer_gather(self.h, local_ptr, row_bytes, height, block_rows, frame_ptr, root))

    def render_rows_image(self, params: GiRenderParams, block_rows, s0, s1, root=0):
        """gi_render_rows_image: this rank's rows rendered, resolved and gathered; the root gets the 8-bit frame [h * w, 3] (others None)."""
        info = self.comm_info()
        rgb = np.empty((params.height * params.width, 3), dtype=np.uint8) if info["rank"] == root else None
        st = GiStats()
        self._ck(self.L.gi_render_rows_image(self.h, C.byref(params), block_rows, s0, s1, _p(rgb), root, C.byref(st)))
        return rgb, st

    def render_tile_dev(self, params: GiRenderParams, x0, y0, x1, y1, s0, s1, accum_ptr):
        st = GiStats()
        self._ck(self.L.gi_render_tile_dev(self.h, C.byref(params), x0, y0, x1, y1, s0, s1, accum_ptr, C.byref(st)))
        return st

    def resolve(self, accum, spp):
        accum = _f64(accum, 3)
        out = np.empty((accum.shape[0], 3), dtype=np.uint8)
        self._ck(self.L.gi_resolve(self.h, accum.shape[0], _p(accum), spp, _p(out)))
        return out

    def resolve_dev(self, n_pixels, accum_ptr, spp, rgb8_ptr):
        self._ck(self.L.gi_resolve_dev(self.h, n_pixels, accum_ptr, spp, rgb8_ptr))
