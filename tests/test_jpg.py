"""JPEG without Qt (csrc/host/gi_jpg.cpp): the decoder that feeds imageTexture (material.h:51-81) for the other
format QImage reads.  A JPEG file's pixels are defined by the decoder's inverse DCT, chroma upsampling and colour conversion; Qt uses
libjpeg with its defaults, and so does PIL — an independent libjpeg build — so every flavour is compared with it BYTE FOR BYTE:
baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 / 4:1:1, grey, odd sizes (edge blocks, edge columns of the triangle filter), restart
intervals, optimised tables, the reference's own JPG asset.  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

PIL = pytest.importorskip("PIL.Image")


def _pil_rgba(path):
    return np.asarray(PIL.open(path).convert("RGBA"))


def _picture(w, h, seed, grey=False):
    """smooth gradients + edges + noise: exercises every DCT coefficient and the chroma filters"""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([127 + 120 * np.sin(x / 7.0 + seed) * np.cos(y / 5.0), 255.0 * ((x // 9 + y // 6) % 2), 40 + 200.0 * x / max(w - 1, 1)], axis=2)
    a += rng.randint(-30, 31, a.shape)
    a = np.clip(a, 0, 255).astype(np.uint8)
    return a[:, :, 0] if grey else a


CASES = [  # name, size, save options
    ("444", (37, 29), dict(subsampling=0, quality=90)),
    ("422", (37, 29), dict(subsampling=1, quality=85)),
    ("420", (37, 29), dict(subsampling=2, quality=75)),
    ("420_even", (64, 48), dict(subsampling=2, quality=95)),
    ("420_tiny", (3, 5), dict(subsampling=2, quality=80)),
    ("420_narrow", (5, 40), dict(subsampling=2, quality=80)),      # 3 chroma columns: the triangle filter's smallest case
    ("420_2cols", (4, 9), dict(subsampling=2, quality=80)),        # 2 chroma columns: box replication instead
    ("422_prog", (53, 31), dict(subsampling=1, quality=80, progressive=True)),
    ("420_prog", (131, 77), dict(subsampling=2, quality=60, progressive=True)),
    ("444_prog_q100", (40, 40), dict(subsampling=0, quality=100, progressive=True)),
    ("420_opt", (90, 61), dict(subsampling=2, quality=50, optimize=True)),
    ("420_lowq", (120, 80), dict(subsampling=2, quality=8)),
    ("444_q100", (33, 17), dict(subsampling=0, quality=100)),
    ("grey", (41, 23), dict(quality=85)),
    ("grey_prog", (70, 45), dict(quality=70, progressive=True)),
]


@pytest.mark.parametrize("name,size,opts", CASES, ids=[c[0] for c in CASES])
def test_decode_matches_libjpeg(lib_built, tmp_path, name, size, opts):
    from gi_raytracer_b200 import host
    w, h = size
    a = _picture(w, h, seed=len(name) * 13 + w, grey=name.startswith("grey"))
    p = str(tmp_path / f"{name}.jpg")
    PIL.fromarray(a).save(p, "JPEG", **opts)
    got = host.jpg_decode(p)
    want = _pil_rgba(p)
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{name}: {(got != want).any(axis=2).sum()} of {w * h} pixels differ, max {np.abs(got.astype(int) - want.astype(int)).max()}"


def test_restart_intervals_and_sampling_440_411(lib_built, tmp_path):
    """flavours PIL's writer only produces through libjpeg options: restart markers, vertical-only (4:4:0) and 4:1:1 subsampling"""
    from gi_raytracer_b200 import host
    a = _picture(75, 58, seed=5)
    for i, kw in enumerate([dict(restart_marker_blocks=3), dict(restart_marker_rows=1), dict(restart_marker_blocks=1, progressive=True)]):
        p = str(tmp_path / f"rst{i}.jpg")
        try:
            PIL.fromarray(a).save(p, "JPEG", quality=80, subsampling=2, **kw)
        except TypeError:
            pytest.skip("this PIL cannot write restart markers")
        if b"\xff\xd0" not in open(p, "rb").read():
            pytest.skip("this PIL ignores the restart options")
        assert np.array_equal(host.jpg_decode(p), _pil_rgba(p)), kw
    for sub in ("4:4:0", "4:1:1"):
        p = str(tmp_path / f"s{sub.replace(':', '')}.jpg")
        try:
            PIL.fromarray(a).save(p, "JPEG", quality=85, subsampling=sub)
        except Exception:
            continue
        assert np.array_equal(host.jpg_decode(p), _pil_rgba(p)), sub


def test_reference_jpg_assets(lib_built):
    """every JPG among the reference's scene assets (where mounted / staged) decodes to libjpeg's pixels"""
    from gi_raytracer_b200 import host
    files = sorted(set(glob.glob("/root/reference/scenes/**/*.jpg", recursive=True) + glob.glob("/root/reference/scenes/**/*.JPG", recursive=True)
                       + glob.glob(os.path.join(ROOT, "scenes", "_assets", "**", "*.jpg"), recursive=True)))
    if not files:
        pytest.skip("no JPG assets present")
    for f in files:
        assert np.array_equal(host.jpg_decode(f), _pil_rgba(f)), f


def test_rejects_what_it_does_not_decode(lib_built, tmp_path):
    from gi_raytracer_b200 import host
    p = str(tmp_path / "not.jpg")
    open(p, "wb").write(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(ValueError):
        host.jpg_decode(p)
    q = str(tmp_path / "cut.jpg")
    PIL.fromarray(_picture(40, 30, 1)).save(q, "JPEG")
    data = open(q, "rb").read()
    open(q, "wb").write(data[:len(data) // 3])   # truncated inside the scan: no EOI
    with pytest.raises(ValueError):
        host.jpg_decode(q)
    c = str(tmp_path / "cmyk.jpg")
    PIL.fromarray(np.zeros((8, 8, 4), np.uint8), "CMYK").save(c, "JPEG")
    with pytest.raises(ValueError):
        host.jpg_decode(c)


def test_image_texture_loads_a_jpg(lib_built, tmp_path):
    """imTex in a scene file names a .jpg: the loader decodes it itself (material.h:57), no sidecar; a JPEG has no alpha channel"""
    from gi_raytracer_b200 import host
    a = _picture(32, 16, seed=9)
    PIL.fromarray(a).save(str(tmp_path / "wall.jpg"), "JPEG", quality=90, subsampling=2)
    (tmp_path / "q.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1\n")
    (tmp_path / "s.scn").write_text("imTex wall.jpg 2 3\ncolorTex 0 0 0\nmat 0 1 1 1 1\nmesh q.obj 0 0 0 0 0 0 0\n")
    sc = host.load_scene(str(tmp_path / "s.scn"))
    t = sc.tex[0]
    assert (t["width"], t["height"], t["has_alpha"], t["tile_u"], t["tile_v"]) == (32, 16, 0, 2.0, 3.0)
    want = _pil_rgba(str(tmp_path / "wall.jpg"))
    assert np.array_equal(sc.tex_pixels[int(t["pixel_offset"]):int(t["pixel_offset"]) + 32 * 16 * 4].reshape(16, 32, 4), want)
