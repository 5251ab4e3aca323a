#!/usr/bin/env python3
"""A/B of the frame schedule (gi_ctx::sched_mode, stream priorities) in ONE process: for every scene the frame time (median of N
frames after two warm-up frames) and the SHA-1 of the fp64 accumulator under each variant — the accumulators must be byte-identical.
usage: python profiles/sched_ab.py [--scenes caustics,glass,...] [--frames 5] [--variants 0:0,1:1,1:1:8:65536]   (sched_mode:GI_STREAM_PRIO[:ring[:tail_threshold]])"""
import argparse, hashlib, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gi_raytracer_b200 import host
from gi_raytracer_b200.abi import render_params
from gi_raytracer_b200.capi import Context
import torch

CASES = {   # scene: (w, h, spp, depth, photons, photon depth)
    "cornell": (512, 512, 16, 4, 750000, 5),
    "caustics": (1024, 1024, 8, 64, 1000000, 5),
    "glass": (1920, 1080, 8, 64, 275000, 5),
    "foliage": (1920, 1080, 4, 64, 0, 5),
    "sponza": (3840, 2160, 1, 64, 0, 5),
}
ap = argparse.ArgumentParser()
ap.add_argument("--scenes", default="caustics,cornell,glass,foliage,sponza")
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--variants", default="0:0,1:1,2:1")
ap.add_argument("--serial", action="store_true", help="also render one frame on ONE stream (overlap off) and compare its bytes")
a = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for scene in a.scenes.split(","):
    w, h, spp, depth, photons, pdepth = CASES[scene]
    sc = host.load_scene(os.path.join(root, "scenes", scene, scene + ".scn"))
    P = render_params(w, h, spp, max_depth=depth, seed=1)
    acc = torch.zeros((w * h, 3), dtype=torch.float64, device="cuda")
    ref = None
    for var in a.variants.split(","):
        f4 = [int(x) for x in var.split(":")] + [8, 32768]
        mode, prio = f4[0], f4[1]
        ring, tail = (f4[2], f4[3]) if len(f4) >= 6 else ((f4[2], 32768) if len(f4) == 5 else (8, 32768))
        os.environ["GI_STREAM_PRIO"] = str(prio)   # read at gi_create
        ctx = Context(0); ctx.upload_scene(sc)
        if photons:
            ctx.photon_trace(photons, pdepth, seed=1); ctx.photon_map_build(None)
        ctx.configure("sched_mode", mode); ctx.configure("ring", ring); ctx.configure("tail_threshold", tail)
        rows = []
        for f in range(a.frames + 2):
            st = ctx.render_tile_dev(P, 0, 0, w, h, 0, spp, acc.data_ptr())
            if f >= 2: rows.append(st)
        med = lambda k: statistics.median(getattr(s, k) for s in rows)
        digest = hashlib.sha1(acc.cpu().numpy().tobytes()).hexdigest()[:16]
        if ref is None: ref = digest
        line = (f"{scene:9s} sched {mode} prio {prio} ring {ring} tail {tail:6d}: frame {med('total_ms'):9.3f} ms (min {min(s.total_ms for s in rows):9.3f}) | bounce {med('trace_ms'):8.3f} direct {med('shadow_ms'):8.3f} "
                f"gather {med('gather_ms'):8.3f} tail {med('shade_ms'):7.3f} | launches {rows[-1].kernel_launches} rays {rows[-1].closest_rays + rows[-1].shadow_rays} sha1 {digest} {'same' if digest == ref else 'DIFFERENT'}")
        print(line, flush=True)
        if a.serial:
            ctx.configure("overlap_threshold", 0)
            ctx.render_tile_dev(P, 0, 0, w, h, 0, spp, acc.data_ptr())
            d2 = hashlib.sha1(acc.cpu().numpy().tobytes()).hexdigest()[:16]
            print(f"{scene:9s} one stream: sha1 {d2} {'same' if d2 == ref else 'DIFFERENT'}", flush=True)
        ctx.close()
