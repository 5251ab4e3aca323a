#!/usr/bin/env python3
"""The five BASELINE configs at their own resolution on one GPU: photon phase, frame time, Mrays/s, gather rate.
C4/C5 use the seeded stand-ins of scenes/make_standins.py; C5 runs 16 of its 1024 spp per GPU (the full count is what
the 8-GPU sample/tile split is for).  usage: python profiles/configs.py [names...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gi_raytracer_b200 import host
from gi_raytracer_b200.abi import render_params
from gi_raytracer_b200.capi import Context
import torch

CFG = {  # name: scene, w, h, spp, max_depth, photons
    "C1": ("cornell", 512, 512, 16, 4, 750000),
    "C2": ("caustics", 1024, 1024, 8, 64, 1000000),
    "C3": ("glass", 1920, 1080, 64, 64, 275000),
    "C4": ("foliage", 1920, 1080, 256, 64, 0),
    "C5": ("sponza", 3840, 2160, 16, 64, 0),
}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = Context(0)
for name in (sys.argv[1:] or list(CFG)):
    scene, w, h, spp, depth, photons = CFG[name]
    t0 = time.time()
    sc = host.load_scene(os.path.join(root, "scenes", scene, scene + ".scn"))
    t_host = time.time() - t0
    ctx.upload_scene(sc)
    nph, pst = ctx.photon_trace(photons, 5, seed=1)
    ctx.photon_map_build(None)
    pm_ms = ctx.kernel_ms("pm_build")[0]
    P = render_params(w, h, spp, max_depth=depth, seed=1)
    acc = torch.zeros((w * h, 3), dtype=torch.float64, device="cuda")
    best = None
    for rep in range(3):   # frames 1 and 2 try the two bounce-kernel forms, frame 3 runs the faster one
        st = ctx.render_tile_dev(P, 0, 0, w, h, 0, spp, acc.data_ptr())
        if best is None or st.total_ms < best.total_ms:
            best = st
    rays = best.closest_rays + best.shadow_rays
    info = ctx.scene_info()
    print(f"{name} {scene:9s} {w}x{h} {spp:4d} spp depth {depth:2d} | prims {sc.n_prims:7d} nodes {info['n_nodes']:8d} refs {info['n_leaf_refs']:9d} full {info['full']} host load+build {t_host:5.1f} s"
          f" | photons {nph:8d} trace {pst.total_ms:8.1f} ms map {pm_ms:6.1f} ms | frame {best.total_ms:10.2f} ms  rays {rays:12d}  {rays / best.total_ms / 1e3:8.1f} Mrays/s"
          f"  (bounce {best.trace_ms:9.2f} direct {best.shadow_ms:9.2f} gather {best.gather_ms:8.2f} tail {best.shade_ms:8.2f} bin {best.bin_ms:7.2f}) gathers {best.gathers:11d}"
          f" nodes/ray {best.closest_node_tests / max(best.closest_rays, 1):6.1f} prims/ray {best.closest_prim_tests / max(best.closest_rays, 1):6.1f} mean {float(acc.mean()) / spp:.6f}", flush=True)
ctx.close()
