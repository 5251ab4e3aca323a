#!/usr/bin/env python3
"""Stage the reference's mounted scene ASSETS (OBJ meshes, textures) into
scenes/_assets/<scene>/ (git-ignored, travels to the GPU box with the snapshot).

Nothing here is reference source code: meshes are copied byte-for-byte, textures are decoded with PIL into the
raw "GIRT" sidecar (`<name>.rgba`: magic, u32 w, u32 h, u32 has_alpha, RGBA8 rows top-first) that both the
QImage shim of oracle/_ref and this repo's own host loader read (there is no PNG decoder in the image besides PIL).
Usage: stage_assets.py <reference scenes dir> <out dir>
"""
import os
import shutil
import struct
import sys


def decode(src, dst):
    from PIL import Image
    im = Image.open(src)
    has_alpha = im.mode in ("RGBA", "LA") or ("transparency" in im.info)
    rgba = im.convert("RGBA")
    with open(dst, "wb") as f:
        f.write(b"GIRT")
        f.write(struct.pack("<III", rgba.width, rgba.height, 1 if has_alpha else 0))
        f.write(rgba.tobytes())


def main(src_root, out_root):
    n = 0
    for scene in sorted(os.listdir(src_root)):
        sdir = os.path.join(src_root, scene)
        if not os.path.isdir(sdir):
            continue
        odir = os.path.join(out_root, scene)
        os.makedirs(odir, exist_ok=True)
        for fn in sorted(os.listdir(sdir)):
            src = os.path.join(sdir, fn)
            low = fn.lower()
            if low.endswith(".obj"):
                shutil.copyfile(src, os.path.join(odir, fn)); n += 1
            elif low.endswith((".png", ".jpg", ".jpeg")) and not low.startswith("render"):
                decode(src, os.path.join(odir, fn + ".rgba")); n += 1
                if low.endswith(".png"):   # the product's own loader decodes PNG itself (csrc/host/gi_png.cpp); the sidecar stays for the QImage shim and as the PIL cross-check
                    shutil.copyfile(src, os.path.join(odir, fn)); n += 1
    print(f"staged {n} asset files into {out_root}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
