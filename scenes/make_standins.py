#!/usr/bin/env python3
"""Seeded procedural stand-ins for the two BASELINE configs whose assets are not mounted:
  scenes/foliage/  C4: ground + 12 000 alpha-textured cards (grass-blade sized) (stochastic alpha cut-outs, soft shadows from a large area light)
  scenes/sponza/   C5: a Sponza-sized atrium (~270 k triangles: long thin columns, arches, wall panels)
The .scn files are committed; the meshes / texture are generated here (git-ignored, they travel to the GPU box).
usage: python scenes/make_standins.py [--force]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import synth  # noqa: E402


def make(force=False):
    fo, sp = os.path.join(HERE, "foliage"), os.path.join(HERE, "sponza")
    os.makedirs(fo, exist_ok=True); os.makedirs(sp, exist_ok=True)
    if force or not os.path.exists(os.path.join(fo, "cards.obj")):
        synth.terrain(os.path.join(fo, "ground.obj"), n=96, size=16.0, amp=0.25, seed=0x5EED0004 & 0xFFFF)
        synth.cards(os.path.join(fo, "cards.obj"), count=12000, seed=0x5EED0004 & 0xFFFF, area=14.0, h=0.35, wscale=0.35)
        synth.leaf_texture(os.path.join(fo, "leaf.png"), size=256, seed=4)
    if force or not os.path.exists(os.path.join(sp, "atrium.obj")):
        n = synth.atrium(os.path.join(sp, "atrium.obj"), cols=64, seg=128, floors=4, seed=0x5EED0005 & 0xFFFF)
        print("sponza stand-in:", n, "triangles")


if __name__ == "__main__":
    make("--force" in sys.argv)
