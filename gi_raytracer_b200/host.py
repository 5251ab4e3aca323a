"""The host-side scene path (C++ classes in csrc/host, mirrored from the reference's scene API) seen from Python:
load a `.scn` through loadScene, rebuild the octree with the reference's rules, flatten it to `SceneArrays`.
Pure host code — works without a GPU."""
import ctypes as C

import numpy as np

from .abi import GiLight, GiMaterial, GiTexture, SceneArrays
from .capi import load_library


def _np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    dt = np.dtype(dtype)
    buf = (C.c_uint8 * (n * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


def prim_boxes(path, quiet=True):
    """loadScene(path) -> Entity::boundingBox() of every primitive, [n][6] (input of gi_octree_build); host only."""
    L = load_library()
    h = C.c_void_p()
    rc = L.gih_scene_load(str(path).encode(), 1 if quiet else 0, C.byref(h))
    if rc != 0:
        raise RuntimeError(f"gih_scene_load({path}) failed: {rc}")
    try:
        n = L.gih_scene_desc(h).contents.n_prims
        out = np.empty((n, 6))
        if n:
            L.gih_scene_prim_bbox(h, out.ctypes.data)
    finally:
        L.gih_scene_free(h)
    return out


def load_scene(path, quiet=True) -> SceneArrays:
    """loadScene(path) + Octree::rebuild() + Octree::flatten() -> SceneArrays (a deep copy; the C++ objects are freed)."""
    L = load_library()
    h = C.c_void_p()
    rc = L.gih_scene_load(str(path).encode(), 1 if quiet else 0, C.byref(h))
    if rc != 0:
        raise RuntimeError(f"gih_scene_load({path}) failed: {rc}")
    try:
        d = L.gih_scene_desc(h).contents
        nn, nr, npm = d.n_nodes, d.n_refs, d.n_prims
        cam = d.camera
        camera = np.array(list(cam.pos) + list(cam.forward) + list(cam.up) + list(cam.right) + [cam.sensor_diag, cam.focal_dist])
        k = [C.c_int() for _ in range(4)]
        nt = C.c_double()
        L.gih_scene_knobs(h, C.byref(k[0]), C.byref(k[1]), C.byref(k[2]), C.byref(k[3]), C.byref(nt))
        sc = SceneArrays(
            node_box=_np(d.node_box, nn * 6, np.float64), node_child=_np(d.node_child, nn, np.uint32), node_mask=_np(d.node_mask, nn, np.uint8),
            node_prim_off=_np(d.node_prim_off, nn, np.uint32), node_prim_cnt=_np(d.node_prim_cnt, nn, np.uint32), leaf_prims=_np(d.leaf_prims, nr, np.uint32),
            prim_type=_np(d.prim_type, npm, np.uint8), prim_geom=_np(d.prim_geom, npm * 9, np.float64), prim_nrm=_np(d.prim_nrm, npm * 9, np.float64),
            prim_uv=_np(d.prim_uv, npm * 6, np.float64), prim_fnorm=_np(d.prim_fnorm, npm * 3, np.float64), prim_mat=_np(d.prim_mat, npm, np.uint32),
            mats=_np(d.mats, d.n_mats, SceneArrays.MAT_DTYPE), tex=_np(d.tex, d.n_tex, SceneArrays.TEX_DTYPE),
            tex_pixels=_np(d.tex_pixels, d.tex_pixel_bytes, np.uint8), lights=_np(d.lights, d.n_lights * 11, np.float64), camera=camera,
            ambient=np.array(list(d.ambient)),
            fogs=_np(d.fogs, d.n_fog, SceneArrays.FOG_DTYPE), fog_grid=_np(d.fog_grid, d.fog_grid_count, np.float64),
            knobs=dict(photons=k[0].value, photon_depth=k[1].value, min_samples=k[2].value, max_samples=k[3].value, noise_thresh=nt.value))
    finally:
        L.gih_scene_free(h)
    return sc


def png_decode(path):
    """csrc/host/gi_png.cpp: (rgba uint8 [h][w][4], has_alpha) — what QImage gave the reference's imageTexture."""
    L = load_library()
    w, h, a = C.c_int(), C.c_int(), C.c_int()
    if L.gih_png_decode(str(path).encode(), C.byref(w), C.byref(h), C.byref(a), None, 0) != 0:
        raise ValueError(f"not a decodable PNG: {path}")
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    if L.gih_png_decode(str(path).encode(), C.byref(w), C.byref(h), C.byref(a), out.ctypes.data, out.size) != 0:
        raise ValueError(f"not a decodable PNG: {path}")
    return out, bool(a.value)


def jpg_decode(path):
    """csrc/host/gi_jpg.cpp: rgba uint8 [h][w][4] (alpha 255) — libjpeg's pixels, what QImage gave the reference's imageTexture."""
    L = load_library()
    w, h = C.c_int(), C.c_int()
    if L.gih_jpg_decode(str(path).encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise ValueError(f"not a decodable JPEG: {path}")
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    if L.gih_jpg_decode(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data, out.size) != 0:
        raise ValueError(f"not a decodable JPEG: {path}")
    return out


def png_encode(path, rgb):
    """8-bit RGB [h][w][3] -> PNG file."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    if load_library().gih_png_encode(str(path).encode(), rgb.shape[1], rgb.shape[0], rgb.ctypes.data) != 0:
        raise ValueError(f"cannot write {path}")
