"""PNG without Qt (csrc/host/gi_png.cpp): the decoder that feeds imageTexture (material.h:51-81) and the
encoder that saves the frame (gui.h:39-45), checked against PIL — an independent libpng-class decoder — on every PNG
flavour (colour types 0/2/3/4/6, depths 1..16, tRNS, Adam7) and on the reference's own texture files; CPU only."""
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import ROOT

PIL = pytest.importorskip("PIL.Image")


def _pil_rgba(path):
    im = PIL.open(path)
    has_alpha = im.mode in ("RGBA", "LA", "PA") or ("transparency" in im.info)
    return np.asarray(im.convert("RGBA")), has_alpha


@pytest.mark.parametrize("mode,interlace", [("RGB", 0), ("RGBA", 0), ("L", 0), ("LA", 0), ("P", 0), ("1", 0), ("RGB", 1), ("RGBA", 1), ("P", 1), ("L", 1)])
def test_decode_matches_pil(lib_built, tmp_path, mode, interlace):
    from gi_raytracer_b200 import host
    rng = np.random.RandomState(len(mode) * 7 + interlace)
    w, h = 37, 29   # odd sizes: partial bytes at sub-byte depths, short Adam7 passes
    if mode == "1":
        im = PIL.fromarray((rng.rand(h, w) > 0.5).astype(np.uint8) * 255).convert("1")
    elif mode == "P":
        im = PIL.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).quantize(13)
    else:
        ch = {"RGB": 3, "RGBA": 4, "L": 1, "LA": 2}[mode]
        a = rng.randint(0, 256, (h, w, ch), dtype=np.uint8)
        im = PIL.fromarray(a[:, :, 0] if ch == 1 else a, mode)
    p = str(tmp_path / f"t_{mode}_{interlace}.png")
    if interlace:
        # PIL cannot write Adam7: re-encode the scanlines by hand
        _write_adam7(p, im)
    else:
        kw = {"transparency": 3} if mode == "P" else {}
        im.save(p, **kw)
    got, ga = host.png_decode(p)
    want, wa = _pil_rgba(p)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert ga == wa


def _write_adam7(path, im):
    """Adam7-interlaced PNG of a PIL image (8-bit L / RGB / RGBA, or P with its palette), filter 0."""
    mode = im.mode
    a = np.asarray(im)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, ch = a.shape
    ctype = {"L": 0, "RGB": 2, "RGBA": 6, "P": 3}[mode]
    raw = b""
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        sub = a[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        for row in sub:
            raw += b"\x00" + row.astype(np.uint8).tobytes()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 1))
    if mode == "P":
        out += chunk(b"PLTE", bytes(im.getpalette()[:256 * 3]))
    out += chunk(b"IDAT", zlib.compress(raw)[:100]) + chunk(b"IDAT", zlib.compress(raw)[100:]) + chunk(b"IEND", b"")   # two IDAT chunks
    with open(path, "wb") as f:
        f.write(out)


def test_decode_sub_byte_grey_16_bit_and_colour_key(lib_built, tmp_path):
    from gi_raytracer_b200 import host

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    def png(w, h, depth, ctype, rows, extra=b""):
        raw = b"".join(b"\x00" + r for r in rows)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + extra + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")

    p = str(tmp_path / "g2.png")   # 2-bit grey, 5 pixels: 0 1 2 3 1 -> 0 85 170 255 85, value 2 is the colour key
    open(p, "wb").write(png(5, 1, 2, 0, [bytes([0b00011011, 0b01000000])], chunk(b"tRNS", b"\x00\x02")))
    got, a = host.png_decode(p)
    assert got[0, :, 0].tolist() == [0, 85, 170, 255, 85] and got[0, :, 3].tolist() == [255, 255, 0, 255, 255] and a
    want, _ = _pil_rgba(p)
    assert np.array_equal(got[:, :, :3], want[:, :, :3])
    p = str(tmp_path / "rgb16.png")   # 16-bit RGB: the high byte is kept (QColor::red() of a 16-bit channel is value >> 8)
    px = np.array([[0x1234, 0xABCD, 0xFFFF], [0x00FF, 0x0100, 0x8000]], dtype=">u2")
    open(p, "wb").write(png(2, 1, 16, 2, [px.tobytes()]))
    got, a = host.png_decode(p)
    assert got[0].tolist() == [[0x12, 0xAB, 0xFF, 255], [0x00, 0x01, 0x80, 255]] and not a
    # every filter type on real data: PIL chooses filters adaptively for photographic content
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:64, 0:80]
    img = np.stack([(xx * 3 + yy) % 256, (xx + yy * 2) % 256, (xx * yy // 7) % 256], axis=2).astype(np.uint8) ^ (rng.rand(64, 80, 3) > 0.97).astype(np.uint8) * 255
    p = str(tmp_path / "filters.png")
    PIL.fromarray(img).save(p, optimize=True)
    got, _ = host.png_decode(p)
    assert np.array_equal(got[:, :, :3], img)
    # corrupt data is refused, not decoded
    bad = bytearray(open(p, "rb").read()); bad[60] ^= 0xFF
    open(str(tmp_path / "bad.png"), "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        host.png_decode(str(tmp_path / "bad.png"))
    with pytest.raises(ValueError):
        host.png_decode(str(tmp_path / "missing.png"))


def test_encode_round_trip_and_pil_reads_it(lib_built, tmp_path):
    from gi_raytracer_b200 import host
    rng = np.random.RandomState(1)
    for w, h in ((1, 1), (33, 7), (256, 144)):
        rgb = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / f"o_{w}x{h}.png")
        host.png_encode(p, rgb)
        assert np.array_equal(np.asarray(PIL.open(p).convert("RGB")), rgb)   # an independent decoder reads our file
        got, a = host.png_decode(p)
        assert np.array_equal(got[:, :, :3], rgb) and (got[:, :, 3] == 255).all() and not a


@pytest.mark.skipif(not os.path.isdir("/root/reference/scenes"), reason="reference scenes not mounted (build container only)")
@pytest.mark.parametrize("rel", ["glass/sandstone.png", "foliage/T_FieldGrass_01_D.png", "cornell/render.png"])
def test_reference_texture_files_decode_like_pil(lib_built, rel):
    """The PNG textures the reference's scenes name (material.h:57 QImage(fname)) and one of its own renders."""
    from gi_raytracer_b200 import host
    p = os.path.join("/root/reference/scenes", rel)
    got, ga = host.png_decode(p)
    want, wa = _pil_rgba(p)
    assert np.array_equal(got, want) and ga == wa


def test_loader_decodes_png_textures_itself(lib_built, tmp_path):
    """imTex in a scene file: the PNG is decoded by the loader (no sidecar), pixels land in gi_scene_desc.tex_pixels."""
    from gi_raytracer_b200 import host
    rng = np.random.RandomState(4)
    tex = rng.randint(0, 256, (8, 16, 4), dtype=np.uint8)
    PIL.fromarray(tex, "RGBA").save(str(tmp_path / "t.png"))
    (tmp_path / "q.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1\n")
    (tmp_path / "s.scn").write_text("imTex t.png 2 3\ncolorTex 0 0 0\nmat 0 1 1 1 1\nmesh q.obj 0 0 0 0 0 0 0\n")
    sc = host.load_scene(str(tmp_path / "s.scn"))
    t = sc.tex[0]
    assert (t["width"], t["height"], t["has_alpha"], t["tile_u"], t["tile_v"]) == (16, 8, 1, 2.0, 3.0)
    assert np.array_equal(sc.tex_pixels[int(t["pixel_offset"]):int(t["pixel_offset"]) + 16 * 8 * 4].reshape(8, 16, 4), tex)
