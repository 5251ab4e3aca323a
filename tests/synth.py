"""Seeded procedural scenes for tests and for the BASELINE configs whose assets are not mounted with the reference
(C4 foliage, C5 sponza — SURVEY Appendix D).  Everything is written as ordinary `.scn` / OBJ / raw-RGBA files and
loaded through the same host loader as the reference scenes.  No reference data is used here."""
import math
import os
import struct

import numpy as np


def _write_obj(path, verts, norms, uvs, faces):
    with open(path, "w") as f:
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % tuple(v))
        for t in uvs:
            f.write("vt %.6f %.6f\n" % tuple(t))
        for n in norms:
            f.write("vn %.6f %.6f %.6f\n" % tuple(n))
        for a, b, c in faces:
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]))


def terrain(path, n=24, size=8.0, amp=0.35, seed=7, y0=0.0):
    """A bumpy n x n grid in the xz-plane with smooth normals and uvs."""
    rng = np.random.RandomState(seed)
    ph = rng.rand(4) * 6.28
    xs = np.linspace(-size / 2, size / 2, n + 1)
    def hgt(x, z):
        return y0 + amp * (math.sin(1.3 * x + ph[0]) * math.cos(0.9 * z + ph[1]) + 0.5 * math.sin(2.1 * z + ph[2]) * math.sin(1.7 * x + ph[3]))
    verts, norms, uvs = [], [], []
    for j in range(n + 1):
        for i in range(n + 1):
            x, z = xs[i], xs[j]
            e = 1e-3
            dx = (hgt(x + e, z) - hgt(x - e, z)) / (2 * e)
            dz = (hgt(x, z + e) - hgt(x, z - e)) / (2 * e)
            nn = np.array([-dx, 1.0, -dz]); nn /= np.linalg.norm(nn)
            verts.append((x, hgt(x, z), z)); norms.append(tuple(nn)); uvs.append((i / n, j / n))
    faces = []
    idx = lambda i, j: j * (n + 1) + i + 1
    for j in range(n):
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i, j + 1), idx(i + 1, j + 1)
            faces.append(((a, a, a), (c, c, c), (b, b, b)))
            faces.append(((b, b, b), (c, c, c), (d, d, d)))
    _write_obj(path, verts, norms, uvs, faces)
    return 2 * n * n


def uv_sphere(path, center, radius, seg=20, rings=12):
    verts, norms, uvs = [], [], []
    for r in range(rings + 1):
        th = math.pi * r / rings
        for s in range(seg + 1):
            ph = 2 * math.pi * s / seg
            n = (math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph))
            verts.append((center[0] + radius * n[0], center[1] + radius * n[1], center[2] + radius * n[2]))
            norms.append(n); uvs.append((s / seg, r / rings))
    faces = []
    idx = lambda r, s: r * (seg + 1) + s + 1
    for r in range(rings):
        for s in range(seg):
            a, b, c, d = idx(r, s), idx(r, s + 1), idx(r + 1, s), idx(r + 1, s + 1)
            if r > 0:
                faces.append(((a, a, a), (b, b, b), (c, c, c)))
            if r < rings - 1:
                faces.append(((b, b, b), (d, d, d), (c, c, c)))
    _write_obj(path, verts, norms, uvs, faces)
    return len(faces)


def cards(path, count=400, seed=11, area=6.0, h=0.9, y0=0.0, wscale=1.0):
    """Alpha cards: upright quads with random orientation (foliage stand-in); uv covers the full texture."""
    rng = np.random.RandomState(seed)
    verts, norms, uvs, faces = [], [], [(0, 0), (1, 0), (0, 1), (1, 1)], []
    for k in range(count):
        cx, cz = (rng.rand(2) - 0.5) * area
        ang = rng.rand() * math.pi
        w = (0.25 + 0.35 * rng.rand()) * wscale; hh = h * (0.5 + rng.rand())
        dx, dz = math.cos(ang) * w, math.sin(ang) * w
        n = (-math.sin(ang), 0.0, math.cos(ang))
        b = len(verts)
        verts += [(cx - dx, y0, cz - dz), (cx + dx, y0, cz + dz), (cx - dx, y0 + hh, cz - dz), (cx + dx, y0 + hh, cz + dz)]
        norms.append(n)
        ni = len(norms)
        faces.append(((b + 1, 1, ni), (b + 2, 2, ni), (b + 3, 3, ni)))
        faces.append(((b + 3, 3, ni), (b + 2, 2, ni), (b + 4, 4, ni)))
    _write_obj(path, verts, norms, uvs, faces)
    return len(faces)


def leaf_texture(path, size=64, seed=3):
    """RGBA leaf-like texture with alpha holes, written as the raw sidecar the loaders read ("GIRT" format)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    cx, cy = size / 2, size / 2
    r = np.sqrt(((xx - cx) / (size * 0.45)) ** 2 + ((yy - cy) / (size * 0.5)) ** 2)
    alpha = np.where(r < 1.0, 255, 0).astype(np.uint8)
    holes = rng.rand(size, size) < 0.08
    alpha[holes] = 0
    alpha[(r > 0.7) & (r < 1.0) & (rng.rand(size, size) < 0.5)] = 128
    img = np.zeros((size, size, 4), dtype=np.uint8)
    img[..., 0] = 30 + (rng.rand(size, size) * 40).astype(np.uint8)
    img[..., 1] = 110 + (rng.rand(size, size) * 100).astype(np.uint8)
    img[..., 2] = 20 + (rng.rand(size, size) * 30).astype(np.uint8)
    img[..., 3] = alpha
    with open(path + ".rgba", "wb") as f:
        f.write(b"GIRT")
        f.write(struct.pack("<III", size, size, 1))
        f.write(img.tobytes())


def atrium(path, cols=10, seg=24, floors=2, seed=5):
    """A Sponza-like atrium: long thin columns, arches and wall panels (stresses the octree with slivers)."""
    verts, norms, uvs, faces = [], [], [], []
    def quad(p0, p1, p2, p3, n):
        b = len(verts)
        verts.extend([p0, p1, p2, p3]); norms.append(n); ni = len(norms)
        uvs.extend([(0, 0), (1, 0), (0, 1), (1, 1)])
        faces.append(((b + 1, b + 1, ni), (b + 2, b + 2, ni), (b + 3, b + 3, ni)))
        faces.append(((b + 3, b + 3, ni), (b + 2, b + 2, ni), (b + 4, b + 4, ni)))
    L, W, H = 24.0, 10.0, 5.0 * floors
    # floor / ceiling / walls as grids
    g = 16
    for i in range(g):
        for j in range(g):
            x0, x1 = -L / 2 + L * i / g, -L / 2 + L * (i + 1) / g
            z0, z1 = -W / 2 + W * j / g, -W / 2 + W * (j + 1) / g
            quad((x0, 0, z0), (x1, 0, z0), (x0, 0, z1), (x1, 0, z1), (0, 1, 0))
            quad((x0, H, z0), (x0, H, z1), (x1, H, z0), (x1, H, z1), (0, -1, 0))
    for i in range(g):
        x0, x1 = -L / 2 + L * i / g, -L / 2 + L * (i + 1) / g
        for k in range(4 * floors):
            y0, y1 = H * k / (4 * floors), H * (k + 1) / (4 * floors)
            quad((x0, y0, -W / 2), (x0, y1, -W / 2), (x1, y0, -W / 2), (x1, y1, -W / 2), (0, 0, 1))
            quad((x0, y0, W / 2), (x1, y0, W / 2), (x0, y1, W / 2), (x1, y1, W / 2), (0, 0, -1))
    # columns: seg-sided prisms, two rows per floor
    for fl in range(floors):
        for c in range(cols):
            for zrow in (-W / 4, W / 4):
                cx = -L / 2 + L * (c + 0.5) / cols
                y0, y1 = 5.0 * fl, 5.0 * fl + 4.2
                r = 0.22
                for s in range(seg):
                    a0, a1 = 2 * math.pi * s / seg, 2 * math.pi * (s + 1) / seg
                    p0 = (cx + r * math.cos(a0), y0, zrow + r * math.sin(a0)); p1 = (cx + r * math.cos(a1), y0, zrow + r * math.sin(a1))
                    p2 = (cx + r * math.cos(a0), y1, zrow + r * math.sin(a0)); p3 = (cx + r * math.cos(a1), y1, zrow + r * math.sin(a1))
                    am = 0.5 * (a0 + a1)
                    quad(p0, p1, p2, p3, (math.cos(am), 0, math.sin(am)))
                # arch to the next column
                if c + 1 < cols:
                    nx = -L / 2 + L * (c + 1.5) / cols
                    for s in range(seg):
                        t0, t1 = math.pi * s / seg, math.pi * (s + 1) / seg
                        mx, rr = 0.5 * (cx + nx), 0.5 * (nx - cx)
                        q0 = (mx - rr * math.cos(t0), y1 + 0.6 * math.sin(t0), zrow - 0.15); q1 = (mx - rr * math.cos(t1), y1 + 0.6 * math.sin(t1), zrow - 0.15)
                        q2 = (mx - rr * math.cos(t0), y1 + 0.6 * math.sin(t0), zrow + 0.15); q3 = (mx - rr * math.cos(t1), y1 + 0.6 * math.sin(t1), zrow + 0.15)
                        quad(q0, q1, q2, q3, (0, -1, 0))
    _write_obj(path, verts, norms, uvs, faces)
    return len(faces)


MIXED_SCN = """# synthetic test scene: terrain + analytic and meshed primitives, every material kind
samples 4 4 0.0015
photons 4000 5
ambient 0.02 0.03 0.05
camera 6.5 3.2 5.5 0 0.6 0
colorTex 0 0 0
colorTex 0.8 0.8 0.8
colorTex 0.9 0.3 0.2
colorTex 0.95 0.95 0.95
checkerboardTex 0.9 0.9 0.2 0.1 0.1 0.6 8
colorTex 0.3 0.7 0.9
mat 1 0 1 1 1
mat 4 0 1 1 1
mat 3 0 0 1 1
mat 3 0 0 0 1.5
mat 2 0 0.4 1 1
mat 5 0 0.05 0.5 1.33
mesh terrain.obj 0 0 0 0 0 0 1
mesh ball.obj 0 0 0 0 0 0 3
sphere -1.8 1.0 1.2 0.7 2
sphere 2.2 0.9 -1.5 0.6 4
box 1.5 0.9 1.8 0.9 0.9 0.9 0.2 0.5 0.1 0
box -2.5 0.7 -2.0 0.7 0.7 0.7 0 0.3 0 5
light 1.0 6.0 2.0 30 30 30 .15
"""

CARDS_SCN = """# synthetic alpha-card scene (foliage stand-in): ground + textured cards with alpha holes
samples 4 4 0.0012
photons 0 5
ambient 0.018 0.018 0.018
camera 5.0 2.2 4.0 0 0.4 0
colorTex 0 0 0
colorTex 0.5 0.45 0.35
imTex leaf.png 1 1
mat 1 0 1 1 1
mat 2 0 1 1 1
mesh ground.obj 0 0 0 0 0 0 0
mesh cards.obj 0 0 0 0 0 0 1
light -6 9 -4 150 150 150 .5
"""

# like CARDS_SCN with material opacities below 1 on top of the texture alpha (Material::getAlpha = opacity * texture alpha,
# material.h:90-93): a constant-colour ground at opacity 0.6 and the cards at 0.85 — every geometric hit is a stochastic decision
CARDS_OP_SCN = CARDS_SCN.replace("mat 1 0 1 1 1", "mat 1 0 1 0.6 1").replace("mat 2 0 1 1 1", "mat 2 0 1 0.85 1").replace("alpha holes", "alpha holes, opacities below one")

SMALL_SCN = """# twelve-triangle scene: the root is a single leaf (not partitioned)
samples 2 2 0.0015
photons 0 5
colorTex 0 0 0
colorTex 0.7 0.7 0.7
mat 1 0 1 1 1
box 0 0 0 1.5 1.5 1.5 0 0 0 0
light 4 5 3 20 20 20 .1
"""

ATRIUM_SCN = """# synthetic atrium (sponza stand-in, SURVEY Appendix D)
samples 4 4 0.0015
photons 0 5
ambient 0.05 0.05 0.06
camera -10.5 2.0 0.5 6 3.0 -0.5
colorTex 0 0 0
colorTex 0.75 0.7 0.6
mat 1 0 1 1 1
mesh atrium.obj 0 0 0 0 0 0 0
light 0 8.5 0 90 90 90 .3
"""


def _no_keywords_in_comments(text):
    kw = {"imTex", "checkerboardTex", "colorTex", "mat", "multiMat", "mesh", "sphere", "box", "light", "heightFog", "photons", "samples", "ambient", "camera"}
    for line in text.splitlines():
        if line.startswith("#"):
            assert not (set(line.split()) & kw), line


def write_all(d, atrium_cols=10):
    """Write every synthetic scene into directory d; returns {name: scn path}."""
    os.makedirs(d, exist_ok=True)
    out = {}
    terrain(os.path.join(d, "terrain.obj"))
    uv_sphere(os.path.join(d, "ball.obj"), (0.2, 1.3, -0.3), 0.8)
    terrain(os.path.join(d, "ground.obj"), n=12, amp=0.1, seed=9)
    cards(os.path.join(d, "cards.obj"))
    leaf_texture(os.path.join(d, "leaf.png"))
    atrium(os.path.join(d, "atrium.obj"), cols=atrium_cols)
    for name, text in (("mixed", MIXED_SCN), ("cards", CARDS_SCN), ("cards_op", CARDS_OP_SCN), ("small", SMALL_SCN), ("atrium", ATRIUM_SCN)):
        _no_keywords_in_comments(text)
        p = os.path.join(d, name + ".scn")
        with open(p, "w") as f:
            f.write(text)
        out[name] = p
    return out
