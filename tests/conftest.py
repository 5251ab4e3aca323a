import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/gi_ref (the reference compiled from /root/reference)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_cornell():
    return np.load(os.path.join(GOLDEN, "cornell_small.npz"))


@pytest.fixture(scope="session")
def golden_caustics():
    return np.load(os.path.join(GOLDEN, "caustics_small.npz"))


@pytest.fixture(scope="session")
def lib_built():
    """libgi_b200.so, built in-tree (nvcc cross-compiles for sm_100a without a GPU)."""
    from gi_raytracer_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(lib_built):
    """A gi_ctx on cuda:0 — the product path; fails loudly (no CPU fallback) when there is no B200."""
    from gi_raytracer_b200.capi import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def synth_dir(tmp_path_factory):
    import synth
    d = tmp_path_factory.mktemp("synth")
    synth.write_all(str(d))
    return str(d)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and bool((a.view(np.uint8) == b.view(np.uint8)).all())


def scene_path(name):
    return os.path.join(ROOT, "scenes", name, name + ".scn")


def have_assets(name):
    return os.path.isdir(os.path.join(ROOT, "scenes", "_assets", name))
