#!/usr/bin/env python3
"""Latency probe: closest-hit launches of 32 .. 1M rays on the caustics scene (kernel ms from CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gi_raytracer_b200 import host  # noqa: E402
from gi_raytracer_b200.capi import Context  # noqa: E402

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = host.load_scene(os.path.join(root, "scenes", "caustics", "caustics.scn"))
ctx = Context(0)
ctx.upload_scene(sc)
o, d, _ = ctx.camera_rays(1024, 1024, 0, 0, 1024, 1024, 0, 1)
rng = np.random.RandomState(0)
perm = rng.permutation(o.shape[0])
for label, oo, dd in (("coherent", o, d), ("shuffled", o[perm], d[perm])):
    to, td = torch.from_numpy(oo).cuda(), torch.from_numpy(dd).cuda()
    prim = torch.empty(oo.shape[0], dtype=torch.int32, device="cuda")
    for n in (32, 1024, 32768, 1 << 20):  # GI_TRACE_MODE=1 selects the warp-per-ray kernel
        for rep in range(3):
            ctx.trace_closest_dev(n, to.data_ptr(), td.data_ptr(), prim.data_ptr())
            ctx.synchronize()
        ms, _ = ctx.kernel_ms("trace_closest")
        w = ctx.last_work("trace_closest")
        print(f"{label:9s} n={n:8d} {ms:8.4f} ms  {n / ms / 1e3:9.2f} Mrays/s  nodes/ray {w[1] / n:6.1f} prims/ray {w[2] / n:6.1f}")
ctx.close()
