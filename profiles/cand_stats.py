import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from gi_raytracer_b200 import host
from gi_raytracer_b200.capi import Context
sc = host.load_scene('/root/repo/scenes/caustics/caustics.scn')
ctx = Context(0); ctx.upload_scene(sc)
ctx.photon_trace(1000000, 5, seed=1); ctx.photon_map_build(None)
o, d, _ = ctx.camera_rays(1024, 1024, 0, 0, 1024, 1024, 0, 1)
prim, hit, nrm, _ = ctx.trace_closest(o, d)
m = prim != 0xFFFFFFFF
rgb, knn, nc = ctx.gather(hit[m], d[m], 32)
print('queries', nc.size, 'mean', nc.mean(), 'max', nc.max())
for pc in (50, 90, 99, 99.9, 99.99): print(pc, np.percentile(nc, pc))
for t in (256, 512, 768, 1024, 2048, 4096, 16384): print('>', t, int((nc > t).sum()), 'cand share', nc[nc > t].sum() / nc.sum())
print('gather ms', ctx.kernel_ms('gather'))
