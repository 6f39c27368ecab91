"""The drop-in boundary without a GPU: libbtsdsp.so loads, exports exactly what include/btsdsp.h declares,
takes no torch/CUDA types in its signatures, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "btsdsp.h")


@pytest.fixture(scope="module")
def lib_path():
    from openbts_ttsou_b200.build import build
    return build()


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(btsdsp_[a-z0-9_]+)\s*\(", src)))


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "btsdsp.h"\nint main(void){return BTSDSP_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                    "-o", str(tmp_path / "t.o")], check=True)
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in code and "cudaStream" not in code and "#include <cuda" not in code and "float2" not in code


def test_library_exports_every_declared_symbol(lib_path):
    names = declared()
    assert len(names) >= 35
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), "libbtsdsp.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names, "exports and header disagree: %s" % (set(exported) ^ set(names))


def test_python_binding_covers_the_header(lib_path):
    import openbts_ttsou_b200 as pkg
    assert sorted(pkg.EXPORTS) == declared()
    pkg.load_library()


def test_no_cpu_fallback(lib_path):
    import torch
    import openbts_ttsou_b200 as pkg
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only boxes")
    with pytest.raises(pkg.BtsDspError, match="no CUDA device"):
        pkg.BtsDsp(0, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openbts_ttsou_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/gen_lpf_taps.py", ""), "%s mentions the oracle" % f
