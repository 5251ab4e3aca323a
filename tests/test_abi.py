"""The C-ABI library: it builds for sm_100a, loads without a GPU, exports every symbol include/gi_api.h declares, and
refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(lib_built):
    from gi_raytracer_b200 import capi
    L = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "gi_api.h")).read()
    declared = sorted(set(re.findall(r"\b(gi_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(capi.API_SYMBOLS), set(declared) ^ set(capi.API_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.gi_version()


def test_struct_layouts_match_header(lib_built, tmp_path):
    """ctypes mirrors vs the C compiler's view of include/gi_api.h (sizes and a few offsets)."""
    import subprocess
    from gi_raytracer_b200 import abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gi_api.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(gi_texture),sizeof(gi_material),sizeof(gi_light),sizeof(gi_camera),sizeof(gi_render_params),sizeof(gi_stats),sizeof(gi_scene_desc),'
                   'offsetof(gi_scene_desc,camera),offsetof(gi_scene_desc,ambient),offsetof(gi_texture,pixel_offset));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.GiTexture), C.sizeof(abi.GiMaterial), C.sizeof(abi.GiLight), C.sizeof(abi.GiCamera), C.sizeof(abi.GiRenderParams),
            C.sizeof(abi.GiStats), C.sizeof(abi.GiSceneDesc), abi.GiSceneDesc.camera.offset, abi.GiSceneDesc.ambient.offset, abi.GiTexture.pixel_offset.offset]
    assert got == want


def test_no_cpu_fallback_without_device(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gi_raytracer_b200 import capi
    with pytest.raises(capi.GiError) as e:
        capi.Context(0)
    assert e.value.code == -2  # GI_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "gi_raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "gi_oracle" not in text and "liboracle" not in text and "oracle_lib" not in text, os.path.join(dirpath, f)


def test_every_binding_declares_its_argument_types(lib_built):
    """A ctypes function without argtypes passes Python ints as 32-bit C ints: 64-bit pointers and sizes get truncated (a crash on the
    GPU box, nothing on a CPU-only run).  Every declared entry point must have its signature set in capi.load_library."""
    from gi_raytracer_b200 import capi
    L = capi.load_library()
    no_args = {"gi_version"}
    missing = [n for n in capi.API_SYMBOLS if n not in no_args and getattr(L, n).argtypes is None]
    assert not missing, missing
