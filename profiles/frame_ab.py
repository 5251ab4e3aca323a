#!/usr/bin/env python3
"""A/B helper: C2 frame time (median of N frames after warm-up) under the current environment knobs.
usage: [GI_TAIL_THRESHOLD=..] [GI_BIN_THRESHOLD=..] python profiles/frame_ab.py [--scene caustics] [--frames 5]"""
import argparse, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gi_raytracer_b200 import host
from gi_raytracer_b200.abi import render_params
from gi_raytracer_b200.capi import Context
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="caustics"); ap.add_argument("--w", type=int, default=1024); ap.add_argument("--h", type=int, default=1024)
ap.add_argument("--spp", type=int, default=8); ap.add_argument("--depth", type=int, default=64); ap.add_argument("--photons", type=int, default=1000000)
ap.add_argument("--frames", type=int, default=5)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = host.load_scene(os.path.join(root, "scenes", a.scene, a.scene + ".scn"))
ctx = Context(0); ctx.upload_scene(sc)
ctx.photon_trace(a.photons, 5, seed=1); ctx.photon_map_build(None)
P = render_params(a.w, a.h, a.spp, max_depth=a.depth, seed=1)
acc = torch.zeros((a.w * a.h, 3), dtype=torch.float64, device="cuda")
rows = []
for f in range(a.frames + 2):
    st = ctx.render_tile_dev(P, 0, 0, a.w, a.h, 0, a.spp, acc.data_ptr())
    if f >= 2: rows.append(st)
med = lambda k: statistics.median(getattr(s, k) for s in rows)
print(f"frame {med('total_ms'):8.3f} ms | bounce {med('trace_ms'):7.3f} direct {med('shadow_ms'):7.3f} gather {med('gather_ms'):7.3f} tail {med('shade_ms'):7.3f} bin {med('bin_ms'):6.3f} | launches {rows[-1].kernel_launches} rays {rows[-1].closest_rays + rows[-1].shadow_rays} tail rays {rows[-1].tail_closest_rays + rows[-1].tail_shadow_rays} checksum {float(acc.sum()):.12e}")
ctx.close()
