"""pytest configuration: markers, import paths, shared fixtures.

  -m "not gpu" : oracle vs golden vectors, host emulation of the kernels vs oracle, C-ABI load/export
                 checks, world_size-2 gloo sharding -- runs on a CPU-only box in a couple of minutes.
  -m gpu       : the parity tests proper, through the C ABI (libbtsdsp.so) on cuda:0.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bits_equal(a, b):
    """bit-for-bit equality of two arrays of the same dtype/shape (NaN-safe, distinguishes +-0 only by value)"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | ((a != a) & (b != b))))


def assert_same(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = ~((a == b) | ((a != a) & (b != b)))
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError("%s: %d of %d values differ, first at %s: %r vs %r" % (
            what, bad.sum(), bad.size, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])]))


@pytest.fixture(scope="session")
def oracle_port():
    from oracle.oracle import Oracle
    return Oracle("port", sps=1)


@pytest.fixture(scope="session")
def oracle_best():
    """the compiled reference (oracle/_ref travels to the GPU box); the C port only stands in for CPU-side tests on a
    box that has neither the reference tree nor the prebuilt library"""
    from oracle.oracle import Oracle
    return Oracle("best", sps=1)


@pytest.fixture(autouse=True)
def _gpu_parity_is_against_the_reference_itself(request):
    """A `-m gpu` test compares the CUDA path with the UNMODIFIED reference compiled into oracle/_ref.  Without that
    library the run would only prove "matches our own port", so it fails instead of falling back."""
    if request.node.get_closest_marker("gpu"):
        from oracle.oracle import have_ref
        assert have_ref(), ("oracle/_ref/libref_oracle.so is missing on this box: build it where /root/reference exists "
                            "(python -c 'import __graft_entry__ as g; g.build()'); GPU parity tests do not fall back to the port")


@pytest.fixture(scope="session")
def hostemu():
    """tests/hostemu/libhostemu.so: the kernels' __host__ __device__ bodies compiled for the CPU"""
    import ctypes
    d = os.path.join(ROOT, "tests", "hostemu")
    so, src = os.path.join(d, "libhostemu.so"), os.path.join(d, "hostemu.cu")
    csrc = os.path.join(ROOT, "openbts_ttsou_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h", ".inc"))]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    for name in ("emu_sinc", "emu_sin_lookup", "emu_cos_lookup"):
        getattr(lib, name).restype = ctypes.c_float
        getattr(lib, name).argtypes = [ctypes.c_float]
    lib.emu_setup(1)
    return lib


@pytest.fixture(scope="session")
def dsp():
    """the product on cuda:0 (fails loudly when the library or the GPU is missing)"""
    import openbts_ttsou_b200 as pkg
    from openbts_ttsou_b200.build import build
    build()
    d = pkg.BtsDsp(0, 1)
    yield d
    d.close()


@pytest.fixture(scope="session")
def dsp4():
    import openbts_ttsou_b200 as pkg
    d = pkg.BtsDsp(0, 4)
    yield d
    d.close()
