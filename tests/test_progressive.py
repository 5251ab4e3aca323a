"""Cancellation and progressive display (second half): RayTracer::stop / start and the `_running` poll of the
row loop (raytracer.h:98, 723-725), the viewer's use of them from another thread (viewer.h:29-62).
Device side: gi_cancel raises a flag that the render / photon calls poll at launch boundaries.  Host side: RayTracer::run
renders bands of rows top to bottom and publishes each finished band; a band that has not started when stop() arrives is
skipped (the reference skips rows), the band on the device is cancelled and not published."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

from conftest import bits_equal, have_assets, scene_path
from gi_raytracer_b200.abi import render_params

pytestmark = pytest.mark.gpu
GI_ERR_CANCELLED = -7


def test_cancel_flag_semantics(ctx, synth_dir):
    from gi_raytracer_b200 import host
    from gi_raytracer_b200.capi import GiError
    sc = host.load_scene(os.path.join(synth_dir, "mixed.scn"))
    ctx.upload_scene(sc)
    ctx.photon_trace(2000, 5, seed=3)
    ctx.photon_map_build(None)
    P = render_params(64, 64, 4, max_depth=6, seed=5)
    full, _ = ctx.render_tile(P, 0, 0, 64, 64, 0, 4)
    ctx.cancel(True)
    try:
        for call in (lambda: ctx.render_tile(P, 0, 0, 64, 64, 0, 4), lambda: ctx.photon_trace(100, 5, seed=3),
                     lambda: ctx.render_adaptive(P, 2, 4, 0.001, 0, 0, 64, 64)):
            with pytest.raises(GiError) as e:   # raised flag: every render / photon call gives up at once, and stays that way
                call()
            assert e.value.code == GI_ERR_CANCELLED
    finally:
        ctx.cancel(False)
    ctx.photon_trace(2000, 5, seed=3)                    # the cancelled photon call left no photons: redo the phase
    ctx.photon_map_build(None)
    again, _ = ctx.render_tile(P, 0, 0, 64, 64, 0, 4)   # lowered: same frame as before, bit for bit
    assert bits_equal(full, again)


@pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")
def test_cancel_from_another_thread_stops_a_render_in_flight(ctx):
    from gi_raytracer_b200 import host
    from gi_raytracer_b200.capi import GiError
    sc = host.load_scene(scene_path("caustics"))
    ctx.upload_scene(sc)
    ctx.photon_trace(20000, 5, seed=3)
    ctx.photon_map_build(None)
    P = render_params(1024, 1024, 32, max_depth=64, seed=5)    # 4 chunks of 2^23 paths, ~65 bounce depths each: > 100 ms
    ctx.render_tile(render_params(256, 256, 1, max_depth=4, seed=5), 0, 0, 256, 256, 0, 1)   # warm-up (buffers, autotune)
    ctx.render_tile(P, 0, 0, 1024, 1024, 0, 32)   # first full-size call: allocations
    t0 = time.time()
    ctx.render_tile(P, 0, 0, 1024, 1024, 0, 32)
    t_full = time.time() - t0
    out = {}

    def work():
        try:
            ctx.render_tile(P, 0, 0, 1024, 1024, 0, 32)
            out["code"] = 0
        except GiError as e:
            out["code"] = e.code
        out["t"] = time.time()

    th = threading.Thread(target=work)
    t0 = time.time()
    th.start()
    time.sleep(0.15 * t_full)
    t_cancel = time.time()
    ctx.cancel(True)
    th.join()
    ctx.cancel(False)
    assert out["code"] == GI_ERR_CANCELLED
    latency = out["t"] - t_cancel
    print(f"full frame {t_full * 1e3:.0f} ms; cancelled after {(t_cancel - t0) * 1e3:.0f} ms, call returned {latency * 1e3:.1f} ms later")
    assert latency < 0.5 * t_full      # it did not run to the end
    small = render_params(64, 64, 2, max_depth=6, seed=5)
    a, _ = ctx.render_tile(small, 0, 0, 64, 64, 0, 2)          # the context is usable afterwards
    b, _ = ctx.render_tile(small, 0, 0, 64, 64, 0, 2)
    assert bits_equal(a, b) and a.max() > 0


def _progressive(L, path, w, h, spp, photons, rows, stop_after):
    rgb = np.zeros((h, w, 3), dtype=np.uint8)
    done, lat = C.c_int(), C.c_double()
    rc = L.gih_render_progressive(path.encode(), w, h, 0, 6, spp, photons, 3, rows, stop_after, rgb.ctypes.data, C.byref(done), C.byref(lat))
    return rc, rgb, done.value, lat.value


def test_progressive_bands_and_stop_from_viewer_thread(lib_built, synth_dir):
    from gi_raytracer_b200 import capi
    L = capi.load_library()
    path = os.path.join(synth_dir, "mixed.scn")
    w, h, spp = 192, 400, 16   # 50 bands of 8 rows: the stop below arrives with ~45 bands still to go
    rc, whole, done, _ = _progressive(L, path, w, h, spp, 2000, 0, -1)        # one device call for the frame
    assert rc == 0 and done == h and whole.max() > 0
    rc, banded, done, _ = _progressive(L, path, w, h, spp, 2000, 24, -1)      # 17 bands (the last one short)
    assert rc == 0 and done == h
    assert np.array_equal(banded, whole)                                     # bands are tiles: identical pixels
    rc, part, done, lat = _progressive(L, path, w, h, spp, 2000, 8, 24)       # stop() once 24 rows are on screen
    print(f"stopped with {done} of {h} rows published, run() returned {lat:.2f} ms after stop()")
    assert rc == 0 and 24 <= done < h and done % 8 == 0
    assert np.array_equal(part[:done], whole[:done])                         # what was published is final
    assert part[done:].max() == 0                                            # rows never started stay cleared (Image::clear, image.h:23)


def test_render_image_equals_render_tile_plus_resolve(ctx, synth_dir):
    """gi_render_image (what RayTracer::run calls per band) = gi_render_tile followed by gi_resolve, without the host round trip."""
    from gi_raytracer_b200 import host
    sc = host.load_scene(os.path.join(synth_dir, "mixed.scn"))
    ctx.upload_scene(sc)
    ctx.photon_trace(2000, 5, seed=3)
    ctx.photon_map_build(None)
    P = render_params(80, 56, 6, max_depth=6, seed=5)
    acc, _ = ctx.render_tile(P, 8, 4, 72, 52, 0, 6)
    img = ctx.resolve(acc, 6)
    rgb, acc2, st = ctx.render_image(P, 8, 4, 72, 52, 0, 6, want_accum=True)
    assert np.array_equal(rgb, img) and bits_equal(acc, acc2) and st.closest_rays > 0
    rgb3, none, _ = ctx.render_image(P, 8, 4, 72, 52, 0, 6)
    assert none is None and np.array_equal(rgb3, img)
