"""Build libgi_b200.so (CUDA kernels + C ABI + the host-side C++ scene classes) and the headless global-illu CLI,
in-tree, for sm_100a.  nvcc cross-compiles without a GPU.  Usage: python -m gi_raytracer_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgi_b200.so")
CLI = os.path.join(HERE, "global-illu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU = [os.path.join(CSRC, "gi_api.cu")]
CPP = [os.path.join(CSRC, "host", f) for f in ("gi_scene.cpp", "gi_loader.cpp", "gi_raytracer.cpp", "gi_host_capi.cpp", "gi_png.cpp", "gi_jpg.cpp")]
HDR = [os.path.join(CSRC, f) for f in ("gi_device.cuh", "gi_kernels.cuh", "gi_octree_build.cuh", os.path.join("host", "api_scene.inc"))] + [os.path.join(CSRC, "host", "gi_scene.hpp"),
                                                                              os.path.join(HERE, "..", "include", "gi_api.h")]
EXTRA = os.environ.get("GI_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wno-unused-result", "-ccbin", "/usr/bin/g++"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(out, extra):
    """An A/B variant of the library with extra nvcc flags (e.g. ['-DGI_REGTOP=0']), written to `out`."""
    subprocess.check_call([NVCC] + list(extra) + FLAGS + ["-shared", "-o", out] + CU + CPP + ["-lz"])
    return out


def build(force=False, verbose=False):
    if force or _stale(LIB, CU + CPP + HDR):
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + CU + CPP + ["-lz"]
        subprocess.check_call(cmd)
    if force or _stale(CLI, [LIB, os.path.join(CSRC, "cli", "global_illu.cpp")]):
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(CSRC, "cli", "global_illu.cpp"), "-o", CLI, "-L" + HERE, "-lgi_b200",
               "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
