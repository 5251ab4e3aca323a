import sys, numpy as np
sys.path.insert(0, '/root/repo')
from gi_raytracer_b200 import host
from gi_raytracer_b200.capi import Context
for scene in ('caustics', 'cornell', 'glass'):
    sc = host.load_scene(f'/root/repo/scenes/{scene}/{scene}.scn')
    ctx = Context(0); ctx.upload_scene(sc)
    n, st = ctx.photon_trace(100000, 5, seed=1)
    print(scene, 'stored', n, 'tries', st.photon_tries, 'traces', st.closest_rays, 'nodes/trace', st.closest_node_tests / st.closest_rays, 'prims/trace', st.closest_prim_tests / st.closest_rays, 'ms', st.total_ms, 'lights', sc.lights)
    ctx.close()
