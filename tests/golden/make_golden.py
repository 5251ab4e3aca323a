#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the reference itself (oracle/_ref/gi_ref = the unmodified reference sources
compiled by oracle/Makefile).  Run in the build container (needs /root/reference):

    make -C oracle ref assets && python tests/golden/make_golden.py

Fixtures (all raw outputs of the reference's own functions, OMP_NUM_THREADS=1, time() interposed):
  cornell_small.npz  scenes/cornell/cornell.scn at 48x48, s in [0,2): scene dump (entities, octree), Halton KATs,
                     sampler KATs, camera rays, closest hits, shadow rays + visibility, 3000 photons, photon-map
                     cells, gather candidates / 32-nearest sets / radiance estimates.
  caustics_small.npz scenes/caustics/caustics.scn at 40x40: rays, hits, shadows, 2500 photons + gather; Octree::intersectSorted's leaf lists
                     (boxes + entry distances) of the primary rays and Octree::intersect's entity lists of the shadow rays.
  api_small.npz      tests/synth.py `small.scn` + the API-built primitives of csrc/host/api_scene.inc (`gi_ref --api-scene 1`: analytic
                     sphere and cones, sphereMesh / coneMesh / quadMesh / boxMesh generators, a checkerboard material) at 56x56, s in
                     [0,2): rays, closest hits (ids, points, normals, uvs), shadow rays + visibility, texture::get / getAlpha at
                     the hits.  Pins sphere::intersect, cone::intersect and checkerboard::get (entities.h:60-101, 158-258; material.h:32-49).
  cones_small.npz    the same base scene + `--api-scene 2`: two analytic cones and a sphere in an UNPARTITIONED root (15 entities) — the only
                     configuration in which the reference can hit a cone at all (entities.h:38-41); pins cone::intersect's hits.
  cards_small.npz    tests/synth.py `cards_op.scn` (alpha-textured cards at opacity 0.85 x texture alpha, ground at opacity 0.6) at 56x56,
                     s in [0,2), run with OMP_NUM_THREADS=1 and time() = 424242: the reference's own xorshift64* stream is then known, and
                     the port replays it (go_trace_closest_replay): ids / hits / shadow bits of the STOCHASTIC alpha path,
                     imageTexture::get / getAlpha (raytracer.h:455, :297; material.h:51-81, 90-93).
  mixed_small.npz    tests/synth.py `mixed.scn` the same way: analytic spheres inside a PARTITIONED octree, a checkerboard, a refractive
                     material with opacity 0.5 (IOR != 1 passes the alpha test whatever the draw).
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


# Tier T6 on CONVERGED images: the five configs' scenes at 40x40, 256 spp, the config's MAX_DEPTH, rendered twice by the
# reference (radiance() per pixel through the row loop of run(), two time() seeds).  Minutes of CPU per scene: run on its own with
#     python tests/golden/make_golden.py converged
CONVERGED = {   # name: scene file stem, directory, max depth, photons
    "C1": ("cornell", "cornell", 4, 100000), "C2": ("caustics", "caustics", 64, 200000), "C3": ("glass", "glass", 64, 100000),
    "C4": ("foliage", "foliage", 64, 0), "C5": ("sponza", "sponza", 64, 0),
}
CONVERGED_RES, CONVERGED_SPP = 40, 256


def converged(only=None):
    root = os.path.dirname(os.path.dirname(HERE))
    for name, (stem, sdir, depth, photons) in CONVERGED.items():
        if only and name not in only:
            continue
        out = {}
        for tag, tv in (("a", 1001), ("b", 2002)):
            d, meta = R.run_ref(os.path.join(root, "scenes", sdir, stem + ".scn"), ["radiance"], threads=os.cpu_count(), time_value=tv, w=CONVERGED_RES, h=CONVERGED_RES,
                                samples=CONVERGED_SPP, max_depth=depth, photons=photons)
            assert meta["radiance_spp"] == CONVERGED_SPP
            out["radiance_" + tag] = R.load(d, "radiance.f64").reshape(-1, 3)
        out["meta_res_spp_depth_photons"] = np.array([CONVERGED_RES, CONVERGED_SPP, depth, photons])
        np.savez_compressed(os.path.join(HERE, f"converged_{name}.npz"), **out)
        print("converged", name, float(out["radiance_a"].mean()), float(out["radiance_b"].mean()), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "converged":
        return converged(sys.argv[2:])
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
    d, meta = R.run_ref(os.path.join(root, "scenes/caustics/caustics.scn"), ["scene", "primary", "queries", "shadow", "photons", "gather"], w=40, h=40, s0=0, s1=1, photons=2500)
    out = pack(d, SCENE_FILES + RAY_FILES + ["ls_off.u32", "ls_box.f64", "ls_t0.f64", "sc_off.u32", "sc_id.u32"])
    out["meta_w_h_s0_s1"] = np.array([40, 40, 0, 1])
    np.savez_compressed(os.path.join(HERE, "caustics_small.npz"), **out)
    print("caustics_small", meta)
    import synth
    sd = tempfile.mkdtemp(prefix="synth_")
    synth.write_all(sd)
    HIT_FILES = ["ray_o.f64", "ray_d.f64", "ray_idx.u32", "hit_id.u32", "hit_pos.f64", "hit_nrm.f64", "hit_uv.f64", "sh_o.f64", "sh_d.f64", "sh_maxt2.f64", "sh_vis.u8",
                 "tex_dif.f64", "tex_em.f64", "tex_alpha.f64"]
    d, meta = R.run_ref(os.path.join(sd, "small.scn"), ["primary", "shadow", "textures"], w=56, h=56, s0=0, s1=2, api_scene=1, photons=0)
    out = pack(d, HIT_FILES)
    out["meta_w_h_s0_s1"] = np.array([56, 56, 0, 2])
    np.savez_compressed(os.path.join(HERE, "api_small.npz"), **out)
    print("api_small", meta)
    d, meta = R.run_ref(os.path.join(sd, "small.scn"), ["primary", "shadow", "textures"], w=56, h=56, s0=0, s1=2, api_scene=2, photons=0)
    out = pack(d, HIT_FILES)
    out["meta_w_h_s0_s1"] = np.array([56, 56, 0, 2])
    np.savez_compressed(os.path.join(HERE, "cones_small.npz"), **out)
    print("cones_small", meta)
    T = 424242
    for scn, name in (("cards_op.scn", "cards_small"), ("mixed.scn", "mixed_small")):
        d, meta = R.run_ref(os.path.join(sd, scn), ["primary", "shadow", "textures"], w=56, h=56, s0=0, s1=2, photons=0, threads=1, time_value=T)
        out = pack(d, HIT_FILES)
        out["meta_w_h_s0_s1"] = np.array([56, 56, 0, 2])
        out["xorshift_seed"] = np.array([T], dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, meta)
    d, meta = R.run_ref(os.path.join(root, "scenes/caustics_fog_dense/caustics_fog_dense.scn"), ["scene", "primary", "fog"], w=40, h=40, s0=0, s1=1, photons=10)
    out = pack(d, ["fog_params.f64", "fog_grid.f64", "fog_pos.f64", "fog_dens.f64", "fog_col.f64", "fogb_hit.u8", "fogb_t.f64"])
    np.savez_compressed(os.path.join(HERE, "fog_small.npz"), **out)
    print("fog_small", meta)


if __name__ == "__main__":
    main()
