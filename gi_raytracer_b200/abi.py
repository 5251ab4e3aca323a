"""ctypes mirror of include/gi_api.h (struct layouts and constants) plus `SceneArrays`, the numpy-side holder of a
flattened scene (the SoA octree / primitive / material / light arrays of `gi_scene_desc`).

This module holds no compute: it only describes memory that crosses the C ABI.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

GI_NO_HIT = 0xFFFFFFFF
GI_PRIM_TRIANGLE, GI_PRIM_SPHERE, GI_PRIM_CONE = 0, 1, 2
GI_TEX_CONST, GI_TEX_CHECKER, GI_TEX_IMAGE = 0, 1, 2


class GiTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tiles", C.c_int32), ("a", C.c_double * 3), ("b", C.c_double * 3),
                ("tile_u", C.c_double), ("tile_v", C.c_double), ("width", C.c_int32), ("height", C.c_int32),
                ("has_alpha", C.c_int32), ("_pad", C.c_int32), ("pixel_offset", C.c_uint64)]


class GiMaterial(C.Structure):
    _fields_ = [("diffuse_tex", C.c_uint32), ("emissive_tex", C.c_uint32), ("roughness", C.c_double),
                ("opacity", C.c_double), ("ior", C.c_double)]


class GiLight(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("col", C.c_double * 3), ("rad", C.c_double), ("dir", C.c_double * 3),
                ("angle", C.c_double)]


class GiFog(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("size", C.c_double * 3), ("col", C.c_double * 3), ("density", C.c_double),
                ("scatter", C.c_double), ("bmin", C.c_double * 3), ("bmax", C.c_double * 3), ("grid_offset", C.c_uint64),
                ("grid_count", C.c_uint64)]


class GiCamera(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("forward", C.c_double * 3), ("up", C.c_double * 3),
                ("right", C.c_double * 3), ("sensor_diag", C.c_double), ("focal_dist", C.c_double)]


class GiSceneDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("node_box", C.c_void_p), ("node_child", C.c_void_p),
                ("node_mask", C.c_void_p), ("node_prim_off", C.c_void_p), ("node_prim_cnt", C.c_void_p),
                ("n_refs", C.c_uint32), ("leaf_prims", C.c_void_p),
                ("n_prims", C.c_uint32), ("prim_type", C.c_void_p), ("prim_geom", C.c_void_p),
                ("prim_nrm", C.c_void_p), ("prim_uv", C.c_void_p), ("prim_fnorm", C.c_void_p),
                ("prim_mat", C.c_void_p),
                ("n_mats", C.c_uint32), ("mats", C.c_void_p), ("n_tex", C.c_uint32), ("tex", C.c_void_p),
                ("tex_pixel_bytes", C.c_uint64), ("tex_pixels", C.c_void_p),
                ("n_lights", C.c_uint32), ("lights", C.c_void_p), ("camera", GiCamera),
                ("ambient", C.c_double * 3),
                ("n_fog", C.c_uint32), ("fogs", C.c_void_p), ("fog_grid_count", C.c_uint64), ("fog_grid", C.c_void_p)]


class GiRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32), ("min_depth", C.c_int32),
                ("spp", C.c_int32), ("k_photons", C.c_int32), ("caustic_max_depth", C.c_int32), ("_pad", C.c_int32),
                ("seed", C.c_uint64)]


class GiStats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("gathers", C.c_uint64),
                ("photon_tries", C.c_uint64), ("photons_stored", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("closest_node_tests", C.c_uint64), ("closest_prim_tests", C.c_uint64), ("shadow_node_tests", C.c_uint64),
                ("shadow_prim_tests", C.c_uint64), ("gather_leaf_depth", C.c_uint64), ("gather_candidates", C.c_uint64),
                ("gather_selected", C.c_uint64),
                ("trace_ms", C.c_double), ("shadow_ms", C.c_double), ("gather_ms", C.c_double),
                ("shade_ms", C.c_double), ("total_ms", C.c_double),
                ("tail_closest_rays", C.c_uint64), ("tail_shadow_rays", C.c_uint64), ("tail_gathers", C.c_uint64),
                ("tail_closest_node_tests", C.c_uint64), ("tail_closest_prim_tests", C.c_uint64), ("tail_shadow_node_tests", C.c_uint64),
                ("tail_shadow_prim_tests", C.c_uint64), ("tail_gather_leaf_depth", C.c_uint64), ("tail_gather_candidates", C.c_uint64),
                ("tail_gather_selected", C.c_uint64), ("bin_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def render_params(width, height, spp, max_depth=64, min_depth=2, k_photons=32, caustic_max_depth=10, seed=1):
    return GiRenderParams(width, height, max_depth, min_depth, spp, k_photons, caustic_max_depth, 0, seed)


def _arr(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


@dataclass
class SceneArrays:
    """A flattened scene in host memory; field meanings are those of gi_scene_desc (include/gi_api.h)."""
    node_box: np.ndarray
    node_child: np.ndarray
    node_mask: np.ndarray
    node_prim_off: np.ndarray
    node_prim_cnt: np.ndarray
    leaf_prims: np.ndarray
    prim_type: np.ndarray
    prim_geom: np.ndarray
    prim_nrm: np.ndarray
    prim_uv: np.ndarray
    prim_fnorm: np.ndarray
    prim_mat: np.ndarray
    mats: np.ndarray           # structured (GiMaterial layout)
    tex: np.ndarray            # structured (GiTexture layout)
    tex_pixels: np.ndarray     # u8
    lights: np.ndarray         # [n][11] f64: pos, col, rad, dir, angle
    camera: np.ndarray         # [14] f64: pos, forward, up, right, sensor_diag, focal_dist
    ambient: np.ndarray        # [3]
    knobs: dict = field(default_factory=dict)   # photons, min_samples, max_samples, noise_thresh (from the .scn)
    fogs: np.ndarray = None    # structured (GiFog layout): heightFog volumes
    fog_grid: np.ndarray = None   # f64 noise grids of all fogs, concatenated

    MAT_DTYPE = np.dtype([("diffuse_tex", "<u4"), ("emissive_tex", "<u4"), ("roughness", "<f8"), ("opacity", "<f8"),
                          ("ior", "<f8")])
    TEX_DTYPE = np.dtype([("kind", "<i4"), ("tiles", "<i4"), ("a", "<f8", 3), ("b", "<f8", 3), ("tile_u", "<f8"),
                          ("tile_v", "<f8"), ("width", "<i4"), ("height", "<i4"), ("has_alpha", "<i4"), ("_pad", "<i4"),
                          ("pixel_offset", "<u8")])

    FOG_DTYPE = np.dtype([("pos", "<f8", 3), ("size", "<f8", 3), ("col", "<f8", 3), ("density", "<f8"), ("scatter", "<f8"),
                          ("bmin", "<f8", 3), ("bmax", "<f8", 3), ("grid_offset", "<u8"), ("grid_count", "<u8")])

    def __post_init__(self):
        assert self.MAT_DTYPE.itemsize == C.sizeof(GiMaterial) and self.TEX_DTYPE.itemsize == C.sizeof(GiTexture)
        assert self.FOG_DTYPE.itemsize == C.sizeof(GiFog)
        self.fogs = np.ascontiguousarray(self.fogs if self.fogs is not None else np.zeros(0, dtype=self.FOG_DTYPE), dtype=self.FOG_DTYPE)
        self.fog_grid = _arr(self.fog_grid if self.fog_grid is not None else np.zeros(0), np.float64)
        self.node_box = _arr(self.node_box, np.float64, (-1, 6))
        self.node_child = _arr(self.node_child, np.uint32)
        self.node_mask = _arr(self.node_mask, np.uint8)
        self.node_prim_off = _arr(self.node_prim_off, np.uint32)
        self.node_prim_cnt = _arr(self.node_prim_cnt, np.uint32)
        self.leaf_prims = _arr(self.leaf_prims, np.uint32)
        self.prim_type = _arr(self.prim_type, np.uint8)
        self.prim_geom = _arr(self.prim_geom, np.float64, (-1, 9))
        self.prim_nrm = _arr(self.prim_nrm, np.float64, (-1, 9))
        self.prim_uv = _arr(self.prim_uv, np.float64, (-1, 6))
        self.prim_fnorm = _arr(self.prim_fnorm, np.float64, (-1, 3))
        self.prim_mat = _arr(self.prim_mat, np.uint32)
        self.mats = np.ascontiguousarray(self.mats, dtype=self.MAT_DTYPE)
        self.tex = np.ascontiguousarray(self.tex, dtype=self.TEX_DTYPE)
        self.tex_pixels = _arr(self.tex_pixels, np.uint8)
        self.lights = _arr(self.lights, np.float64, (-1, 11))
        self.camera = _arr(self.camera, np.float64, (14,))
        self.ambient = _arr(self.ambient, np.float64, (3,))

    @property
    def n_nodes(self):
        return int(self.node_mask.shape[0])

    @property
    def n_prims(self):
        return int(self.prim_type.shape[0])

    @property
    def root_box(self):
        return self.node_box[0].copy()

    def desc(self):
        """A GiSceneDesc pointing at these arrays (which must stay alive while the desc is in use)."""
        d = GiSceneDesc()
        p = lambda a: a.ctypes.data if a.size else None
        d.n_nodes = self.n_nodes
        d.node_box, d.node_child, d.node_mask = p(self.node_box), p(self.node_child), p(self.node_mask)
        d.node_prim_off, d.node_prim_cnt = p(self.node_prim_off), p(self.node_prim_cnt)
        d.n_refs, d.leaf_prims = int(self.leaf_prims.size), p(self.leaf_prims)
        d.n_prims = self.n_prims
        d.prim_type, d.prim_geom, d.prim_nrm = p(self.prim_type), p(self.prim_geom), p(self.prim_nrm)
        d.prim_uv, d.prim_fnorm, d.prim_mat = p(self.prim_uv), p(self.prim_fnorm), p(self.prim_mat)
        d.n_mats, d.mats = int(self.mats.size), p(self.mats)
        d.n_tex, d.tex = int(self.tex.size), p(self.tex)
        d.tex_pixel_bytes, d.tex_pixels = int(self.tex_pixels.size), p(self.tex_pixels)
        d.n_lights, d.lights = int(self.lights.shape[0]), p(self.lights)
        cam = self.camera
        d.camera = GiCamera((C.c_double * 3)(*cam[0:3]), (C.c_double * 3)(*cam[3:6]), (C.c_double * 3)(*cam[6:9]),
                            (C.c_double * 3)(*cam[9:12]), cam[12], cam[13])
        d.ambient = (C.c_double * 3)(*self.ambient)
        d.n_fog, d.fogs = int(self.fogs.size), p(self.fogs)
        d.fog_grid_count, d.fog_grid = int(self.fog_grid.size), p(self.fog_grid)
        return d
