"""The frame schedules give the same frame, bit for bit.  gi_ctx::sched_mode lays a chunk's kernels over the context's three
streams in two ways (0: a depth's shadow rays and gathers behind the next bounce only when the depth is short, everything drained
before the tail; 1: every depth's shadow rays / gathers on the side streams, hit lists in a ring, the tail kernel not waiting for
them); overlap_threshold 0 puts everything on ONE stream.  Per path the sums L / Ld / Lc receive their terms in bounce order
under all of them (raytracer.h:167-276 is one recursion per path), so the fp64 accumulators must be byte-identical — a race
between the streams would show as a difference.  Small tail thresholds make the wavefront run deep enough to wrap the ring."""
import os

import pytest

from conftest import bits_equal, have_assets, scene_path
from gi_raytracer_b200.abi import render_params

pytestmark = pytest.mark.gpu


def _frames(ctx, P, w, h, spp, tail_thresholds):
    out = {}
    try:
        for tt in tail_thresholds:
            ctx.configure("tail_threshold", tt)
            for name, overlap, mode in (("one stream", 0, 0), ("sched 0", 1 << 20, 0), ("sched 1", 1 << 20, 1), ("sched 1, ring of 3", 1 << 20, 13), ("sched 2", 1 << 20, 2),
                                        ("sched 0, long shadow launches only", 1, 0)):
                ctx.configure("overlap_threshold", overlap)
                ctx.configure("sched_mode", mode % 10)
                ctx.configure("ring", 3 if mode >= 10 else 8)
                for rep in range(2):   # twice: the second frame runs with warm buffers and the autotuned bounce form
                    acc, st = ctx.render_tile(P, 0, 0, w, h, 0, spp)
                    out[(tt, name, rep)] = (acc, st)
    finally:
        ctx.configure("tail_threshold", 32768); ctx.configure("overlap_threshold", 1 << 20); ctx.configure("sched_mode", 2); ctx.configure("ring", 8)
    return out


def _check(out, across_thresholds=False):
    t_min = min(k[0] for k in out)
    for k, (acc, st) in out.items():
        k0 = (t_min if across_thresholds else k[0], "one stream", 0)
        a0, s0 = out[k0]
        assert bits_equal(acc, a0), f"frame differs under {k} (reference {k0})"
        assert (st.closest_rays, st.shadow_rays, st.gathers) == (s0.closest_rays, s0.shadow_rays, s0.gathers), f"ray counts differ under {k}"


def test_schedules_bit_identical_synth(ctx, synth_dir):
    from gi_raytracer_b200 import host
    sc = host.load_scene(os.path.join(synth_dir, "mixed.scn"))
    ctx.upload_scene(sc)
    ctx.photon_trace(4000, 5, seed=3)
    ctx.photon_map_build(None)
    P = render_params(96, 96, 4, max_depth=12, seed=5)
    _check(_frames(ctx, P, 96, 96, 4, (0, 64, 32768)), across_thresholds=True)   # 0: no tail at all (every depth a wavefront launch: the ring wraps)


@pytest.mark.skipif(not (have_assets("caustics") and have_assets("glass")), reason="assets not staged")
@pytest.mark.parametrize("scene,photons", [("caustics", 60000), ("glass", 30000)])
def test_schedules_bit_identical_assets(ctx, scene, photons):
    from gi_raytracer_b200 import host
    sc = host.load_scene(scene_path(scene))
    ctx.upload_scene(sc)
    ctx.photon_trace(photons, 5, seed=3)
    ctx.photon_map_build(None)
    P = render_params(320, 240, 8, max_depth=64, seed=7)
    _check(_frames(ctx, P, 320, 240, 8, (16, 2048, 32768)))
