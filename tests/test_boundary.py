"""The drop-in boundary as a user of the reference meets it: C++ written against the reference's scene API compiles
against the mirror (csrc/host/gi_scene.hpp) and — on a B200 — runs; the headless `global-illu` executable renders the same image
as the C ABI called directly."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, have_assets, scene_path

HOST = os.path.join(ROOT, "gi_raytracer_b200", "csrc", "host")
PKG = os.path.join(ROOT, "gi_raytracer_b200")
GLM = "/root/reference/3rd_party"


def _compile(tmp_path, with_glm):
    exe = str(tmp_path / ("boundary_glm" if with_glm else "boundary"))
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", HOST, os.path.join(ROOT, "tests", "boundary_main.cpp"), "-o", exe, "-L" + PKG, "-lgi_b200", "-Wl,-rpath," + PKG]
    if with_glm:
        cmd[1:1] = ["-DGI_TEST_WITH_GLM", "-I", GLM]
    subprocess.check_call(cmd)
    return exe


def test_reference_style_translation_unit_compiles_against_the_mirror(lib_built, tmp_path):
    """main.cpp:30-41's statements + Octree::intersect / intersectSorted + Entity::intersect, with a foreign vec3 type"""
    assert os.path.exists(_compile(tmp_path, False))


@pytest.mark.ref
@pytest.mark.skipif(not os.path.isdir(os.path.join(GLM, "glm")), reason="glm (vendored by the reference) not mounted")
def test_same_translation_unit_with_glm_vectors(lib_built, tmp_path):
    assert os.path.exists(_compile(tmp_path, True))


@pytest.mark.gpu
@pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")
def test_reference_style_program_runs_on_the_gpu(lib_built, tmp_path):
    exe = _compile(tmp_path, False)
    out = subprocess.run([exe, scene_path("caustics"), str(tmp_path / "b.ppm")], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "boundary_main ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")
def test_global_illu_cli_equals_the_c_abi(ctx, tmp_path):
    """the executable (main.cpp's flow on the mirror classes) against gi_render_image called directly: same PNG pixels"""
    from gi_raytracer_b200 import build, host
    from gi_raytracer_b200.abi import render_params
    build.build()
    cli = os.path.join(PKG, "global-illu")
    png = str(tmp_path / "cli.png")
    w = h = 160
    out = subprocess.run([cli, scene_path("caustics"), str(w), str(h), png, "--spp", "3", "--photons", "50000", "--max-depth", "8", "--seed", "4"], capture_output=True, text=True, cwd=ROOT,
                         timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rgba, _ = host.png_decode(png)
    sc = host.load_scene(scene_path("caustics"))
    ctx.upload_scene(sc)
    ctx.photon_trace(50000, 5, seed=4)
    ctx.photon_map_build(None)
    P = render_params(w, h, 3, max_depth=8, seed=4)
    rgb, _, _ = ctx.render_image(P, 0, 0, w, h, 0, 3)
    assert np.array_equal(rgba[..., :3].reshape(-1, 3), rgb)
