#!/usr/bin/env python3
"""Minimal profiling workload: C2 (caustics 1024x1024x8, 1M photons) — photon phase, then `--frames` frames.
Used under ncu (see profiles/README.md); prints per-family kernel times when run plainly."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gi_raytracer_b200 import host  # noqa: E402
from gi_raytracer_b200.abi import render_params  # noqa: E402
from gi_raytracer_b200.capi import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="caustics")
ap.add_argument("--w", type=int, default=1024)
ap.add_argument("--h", type=int, default=1024)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--depth", type=int, default=64)
ap.add_argument("--photons", type=int, default=1000000)
ap.add_argument("--frames", type=int, default=1)
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = host.load_scene(os.path.join(root, "scenes", a.scene, a.scene + ".scn"))
ctx = Context(0)
ctx.upload_scene(sc)
n, st = ctx.photon_trace(a.photons, 5, seed=1)
ctx.photon_map_build(None)
print("photons", n, "trace ms", st.total_ms, "pm build ms", ctx.kernel_ms("pm_build"), ctx.photon_map_info())
P = render_params(a.w, a.h, a.spp, max_depth=a.depth, seed=1)
import torch
acc = torch.zeros((a.w * a.h, 3), dtype=torch.float64, device="cuda")
for f in range(a.frames):
    st = ctx.render_tile_dev(P, 0, 0, a.w, a.h, 0, a.spp, acc.data_ptr())
    print({k: v for k, v in st.as_dict().items()})
ctx.close()
