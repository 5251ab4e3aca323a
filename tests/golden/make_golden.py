#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the reference itself (oracle/_ref/gi_ref = the unmodified reference sources
compiled by oracle/Makefile).  Run in the build container (needs /root/reference):

    make -C oracle ref assets && python tests/golden/make_golden.py

Fixtures (all raw outputs of the reference's own functions, OMP_NUM_THREADS=1, time() interposed):
  cornell_small.npz  scenes/cornell/cornell.scn at 48x48, s in [0,2): scene dump (entities, octree), Halton KATs,
                     sampler KATs, camera rays, closest hits, shadow rays + visibility, 3000 photons, photon-map
                     cells, gather candidates / 32-nearest sets / radiance estimates.
  caustics_small.npz scenes/caustics/caustics.scn at 40x40: rays, hits, shadows, 2500 photons + gather.
  fog_small.npz      scenes/caustics_fog_dense at 40x40 (same geometry, camera and rays as caustics_small): the HeightFog
                     parameters and noise grid as the reference's constructor filled it, Octree::atmosphereDensity at 4000
                     points, Octree::atmosphereBounds on the primary rays.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refdump as R  # noqa: E402

SCENE_FILES = ["ent_type.u8", "ent_pos.f64", "ent_nrm.f64", "ent_uv.f64", "ent_fnorm.f64", "ent_mat.f64", "ent_diftex.u32", "ent_emtex.u32",
               "tex_color.f64", "node_box.f64", "node_mask.u8", "node_cnt.u32", "node_refs.u32", "lights.f64", "camera.f64", "knobs.f64"]
RAY_FILES = ["ray_o.f64", "ray_d.f64", "ray_idx.u32", "hit_id.u32", "hit_pos.f64", "hit_nrm.f64", "hit_uv.f64", "sh_o.f64", "sh_d.f64", "sh_maxt2.f64",
             "sh_vis.u8", "photons.f64", "pm_box.f64", "pm_leaf.u8", "pm_cnt.u32", "pm_refs.u32", "q_pos.f64", "q_dir.f64", "q_est.f64", "q_cand_off.u32",
             "q_cand.u32", "q_knn.u32"]
KAT_FILES = ["halton_idx.u32", "halton_val.f32", "henum_query.u32", "henum_index.u32", "henum_scaled.f32", "kat_in.f64", "kat_out.f64"]


def pack(d, names):
    return {n.replace(".", "_"): R.load(d, n) for n in names}


def main():
    root = os.path.dirname(os.path.dirname(HERE))
    d, meta = R.run_ref(os.path.join(root, "scenes/cornell/cornell.scn"), ["scene", "halton", "samplers", "primary", "shadow", "photons", "gather"],
                        w=48, h=48, s0=0, s1=2, photons=3000)
    out = pack(d, SCENE_FILES + RAY_FILES + KAT_FILES)
    # the Halton table is large (256 dims x 1604 indices): keep every 3rd index
    idx = out["halton_idx_u32"]
    val = out["halton_val_f32"].reshape(256, -1)
    keep = np.arange(0, idx.size, 3)
    out["halton_idx_u32"], out["halton_val_f32"] = idx[keep], val[:, keep].copy()
    out["meta_w_h_s0_s1"] = np.array([48, 48, 0, 2])
    np.savez_compressed(os.path.join(HERE, "cornell_small.npz"), **out)
    print("cornell_small", meta)
    d, meta = R.run_ref(os.path.join(root, "scenes/caustics/caustics.scn"), ["scene", "primary", "shadow", "photons", "gather"], w=40, h=40, s0=0, s1=1, photons=2500)
    out = pack(d, SCENE_FILES + RAY_FILES)
    out["meta_w_h_s0_s1"] = np.array([40, 40, 0, 1])
    np.savez_compressed(os.path.join(HERE, "caustics_small.npz"), **out)
    print("caustics_small", meta)
    d, meta = R.run_ref(os.path.join(root, "scenes/caustics_fog_dense/caustics_fog_dense.scn"), ["scene", "primary", "fog"], w=40, h=40, s0=0, s1=1, photons=10)
    out = pack(d, ["fog_params.f64", "fog_grid.f64", "fog_pos.f64", "fog_dens.f64", "fog_col.f64", "fogb_hit.u8", "fogb_t.f64"])
    np.savez_compressed(os.path.join(HERE, "fog_small.npz"), **out)
    print("fog_small", meta)


if __name__ == "__main__":
    main()
