import sys, os, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from gi_raytracer_b200 import host
from gi_raytracer_b200.capi import Context
from conftest import scene_path
ctx = Context(0)
for name in sys.argv[1:]:
    p = scene_path(name)
    sc = host.load_scene(p); boxes = host.prim_boxes(p)
    for k in range(3):
        if k == 2: os.environ['GI_TRACE_BUILD'] = '1'
        got, ms = ctx.octree_build(sc.prim_type, sc.prim_geom, boxes, sc.root_box)
        print(name, 'build', k, ms, 'ms', got['node_mask'].size, flush=True)
