"""Helpers around oracle/_ref/gi_ref (the reference itself, compiled from /root/reference by oracle/Makefile):
run it, load its raw dumps, and turn its scene dump into a `SceneArrays` (TEST INFRASTRUCTURE)."""
import os
import subprocess
import tempfile

import numpy as np

from gi_raytracer_b200.abi import SceneArrays, GI_TEX_CONST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GI_REF = os.path.join(ROOT, "oracle", "_ref", "gi_ref")
GI_REF_FAST = os.path.join(ROOT, "oracle", "_ref", "gi_ref_fast")
ASSETS = os.path.join(ROOT, "scenes", "_assets")

_EXT = {"f64": np.float64, "f32": np.float32, "u32": np.uint32, "u8": np.uint8}


def have_ref():
    return os.path.exists(GI_REF)


def have_assets(scene="cornell"):
    return os.path.isdir(os.path.join(ASSETS, scene))


def run_ref(scene_path, cmds, outdir=None, threads=1, fast=False, time_value=None, **opts):
    """Run gi_ref; returns (outdir, meta dict)."""
    outdir = outdir or tempfile.mkdtemp(prefix="giref_")
    os.makedirs(outdir, exist_ok=True)
    args = [GI_REF_FAST if fast else GI_REF, scene_path, outdir]
    for k, v in opts.items():
        args += ["--" + k.replace("_", "-"), str(v)]
    args += list(cmds)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if time_value is not None:
        env["GI_REF_TIME"] = str(time_value)
    with open(os.path.join(outdir, "log.txt"), "w") as log:
        subprocess.check_call(args, stdout=log, stderr=subprocess.STDOUT, env=env, cwd=ROOT)
    return outdir, load_meta(outdir)


def load_meta(d):
    meta = {}
    with open(os.path.join(d, "meta.txt")) as f:
        for line in f:
            if "=" in line:
                k, v = line.strip().split("=", 1)
                try:
                    meta[k] = int(v)
                except ValueError:
                    meta[k] = float(v)
    return meta


def load(d, name):
    ext = name.rsplit(".", 1)[1]
    return np.fromfile(os.path.join(d, name), dtype=_EXT[ext])


def preorder_to_bfs(box, mask, cnt, refs):
    """The reference dump lists nodes in DFS pre-order; gi_scene_desc wants breadth-first with contiguous children."""
    n = mask.shape[0]
    popc = np.array([bin(int(m)).count("1") for m in mask], dtype=np.int64)
    # children of each node in pre-order: first child = i+1, next siblings follow the previous sibling's subtree
    size = np.ones(n, dtype=np.int64)
    kids = [[] for _ in range(n)]
    stack = []
    for i in range(n):  # build parent links
        while stack and stack[-1][1] == 0:
            stack.pop()
        if stack:
            kids[stack[-1][0]].append(i)
            stack[-1][1] -= 1
        stack.append([i, int(popc[i])])
    ref_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    order, child = [0], np.zeros(n, dtype=np.uint32)
    newidx = {0: 0}
    q = 0
    while q < len(order):
        i = order[q]
        if kids[i]:
            child[q] = len(order)
            for c in kids[i]:
                newidx[c] = len(order)
                order.append(c)
        q += 1
    order = np.array(order)
    nb = box.reshape(-1, 6)[order]
    nm = mask[order]
    nc = cnt[order]
    off = np.zeros(n, dtype=np.uint32)
    out_refs = []
    pos = 0
    for k, i in enumerate(order):
        off[k] = pos
        out_refs.append(refs[ref_off[i]:ref_off[i + 1]])
        pos += int(cnt[i])
    out_refs = np.concatenate(out_refs) if out_refs else np.zeros(0, dtype=np.uint32)
    return nb, child, nm, off, nc.astype(np.uint32), out_refs.astype(np.uint32)


def scene_from_dump(d):
    """SceneArrays from a `gi_ref ... scene` dump directory."""
    return scene_from_arrays(lambda name: load(d, name))


def scene_from_npz(z):
    """SceneArrays from a tests/golden/*.npz fixture (same arrays, '.' replaced by '_' in the names)."""
    return scene_from_arrays(lambda name: z[name.replace(".", "_")])


def scene_from_arrays(load_):
    """Textures come out as constant colours (the dump records texture::color only), which is exact for scenes
    without imTex/checkerboardTex."""
    d = None
    load = lambda _d, name: load_(name)  # noqa: E731
    box = load(d, "node_box.f64")
    mask = load(d, "node_mask.u8")
    cnt = load(d, "node_cnt.u32")
    refs = load(d, "node_refs.u32")
    nb, child, nm, off, nc, lrefs = preorder_to_bfs(box, mask, cnt, refs)
    matv = load(d, "ent_mat.f64").reshape(-1, 3)
    dif, em = load(d, "ent_diftex.u32"), load(d, "ent_emtex.u32")
    key = np.concatenate([matv, dif[:, None].astype(np.float64), em[:, None].astype(np.float64)], axis=1)
    uniq, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    mats = np.zeros(uniq.shape[0], dtype=SceneArrays.MAT_DTYPE)
    mats["roughness"], mats["opacity"], mats["ior"] = uniq[:, 0], uniq[:, 1], uniq[:, 2]
    mats["diffuse_tex"], mats["emissive_tex"] = uniq[:, 3].astype(np.uint32), uniq[:, 4].astype(np.uint32)
    texcol = load(d, "tex_color.f64").reshape(-1, 3)
    tex = np.zeros(texcol.shape[0], dtype=SceneArrays.TEX_DTYPE)
    tex["kind"] = GI_TEX_CONST
    tex["a"] = texcol
    cam = load(d, "camera.f64")
    knobs = load(d, "knobs.f64")
    fogs, fog_grid = None, None
    try:
        fp = load(d, "fog_params.f64").reshape(-1, 19)
        fog_grid = load(d, "fog_grid.f64")
    except (FileNotFoundError, KeyError):
        fp = np.zeros((0, 19))
    if fp.shape[0]:
        fogs = np.zeros(fp.shape[0], dtype=SceneArrays.FOG_DTYPE)
        fogs["pos"], fogs["size"], fogs["col"], fogs["density"], fogs["scatter"] = fp[:, 0:3], fp[:, 3:6], fp[:, 6:9], fp[:, 9], fp[:, 10]
        fogs["bmin"], fogs["bmax"], fogs["grid_offset"], fogs["grid_count"] = fp[:, 11:14], fp[:, 14:17], fp[:, 17].astype(np.uint64), fp[:, 18].astype(np.uint64)
    return SceneArrays(fogs=fogs, fog_grid=fog_grid, node_box=nb, node_child=child, node_mask=nm, node_prim_off=off, node_prim_cnt=nc, leaf_prims=lrefs,
                       prim_type=load(d, "ent_type.u8"), prim_geom=load(d, "ent_pos.f64"), prim_nrm=load(d, "ent_nrm.f64"),
                       prim_uv=load(d, "ent_uv.f64"), prim_fnorm=load(d, "ent_fnorm.f64"), prim_mat=inv.astype(np.uint32).ravel(),
                       mats=mats, tex=tex, tex_pixels=np.zeros(0, dtype=np.uint8), lights=load(d, "lights.f64"), camera=cam,
                       ambient=knobs[5:8],
                       knobs=dict(photons=int(knobs[0]), min_samples=int(knobs[2]), max_samples=int(knobs[3]), noise_thresh=float(knobs[4])))
