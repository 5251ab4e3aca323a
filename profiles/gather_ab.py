#!/usr/bin/env python3
"""A/B helper: the isolated gather of bench.py (C2: 1 M photons, the 776 666 primary-hit queries) under the library named by GI_LIB.
Prints the time per gather run, the share of each kernel family is in the ncu launch lists; checksums make variants comparable."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gi_raytracer_b200 import host
from gi_raytracer_b200.capi import Context
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = host.load_scene(os.path.join(root, "scenes", "caustics", "caustics.scn"))
ctx = Context(0); ctx.upload_scene(sc)
ctx.photon_trace(1000000, 5, seed=1); ctx.photon_map_build(None)
W = H = 1024
o, d, _ = ctx.camera_rays(W, H, 0, 0, W, H, 0, 1)
prim, hit, nrm, _ = ctx.trace_closest(o, d)
m = prim != 0xFFFFFFFF
nn = nrm[m].copy(); flip = (nn * d[m]).sum(axis=1) > 0; nn[flip] *= -1.0
refl = d[m] - nn * (nn * d[m]).sum(axis=1)[:, None] * 2.0
dev = torch.device("cuda", 0)
q_pos = torch.from_numpy(np.ascontiguousarray(hit[m])).to(dev); q_dir = torch.from_numpy(np.ascontiguousarray(refl)).to(dev)
q_rgb = torch.empty_like(q_pos); nq = q_pos.shape[0]
knn = torch.empty((nq, 32), dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for _ in range(3): ctx.gather_dev(nq, q_pos.data_ptr(), q_dir.data_ptr(), q_rgb.data_ptr(), 32)
ctx.synchronize()
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record(stream)
for _ in range(10): ctx.gather_dev(nq, q_pos.data_ptr(), q_dir.data_ptr(), q_rgb.data_ptr(), 32)
g1.record(stream); ctx.synchronize(); torch.cuda.synchronize()
ms = g0.elapsed_time(g1) / 10
ctx.gather_dev(nq, q_pos.data_ptr(), q_dir.data_ptr(), q_rgb.data_ptr(), 32, knn_ptr=knn.data_ptr()); ctx.synchronize()
h = hashlib.sha1(q_rgb.cpu().numpy().tobytes() + knn.cpu().numpy().tobytes()).hexdigest()[:16]
print(f"gather {ms:7.4f} ms  {nq / ms / 1e3:8.1f} Mq/s  queries {nq}  rgb+knn sha1 {h}")
ctx.close()
