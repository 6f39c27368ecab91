#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Builds tests/cpp/surface_test.cpp against the UNMODIFIED reference sources where they lie
under /root/reference (same flags as oracle/Makefile), runs it, and stores its dump as tests/golden/surface_ref.bin --
the outputs of the reference itself for the parts of the sigProcLib.h surface the main goldens do not reach.
tests/test_gpu_shim.py builds the same program against the btsdsp shim and compares byte for byte.

    python oracle/gen_surface_golden.py [/root/reference]
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SRCS = ["Transceiver/sigProcLib.cpp", "GSM/GSMCommon.cpp", "CommonLibs/BitVector.cpp", "CommonLibs/Logger.cpp",
        "CommonLibs/Threads.cpp", "CommonLibs/Timeval.cpp", "CommonLibs/Sockets.cpp"]


def main():
    if not os.path.isdir(os.path.join(REF, "Transceiver")):
        raise SystemExit("reference tree %s absent: the golden file can only be regenerated where it exists" % REF)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "surface_ref")
        subprocess.run(["g++", "-O3", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w", "-include", "unistd.h",
                        "-I" + os.path.join(REF, "Transceiver"), "-I" + os.path.join(REF, "CommonLibs"),
                        "-I" + os.path.join(REF, "GSM"), os.path.join(ROOT, "tests", "cpp", "surface_test.cpp")] +
                       [os.path.join(REF, s) for s in SRCS] + ["-o", exe], check=True)
        out = os.path.join(ROOT, "tests", "golden", "surface_ref.bin")
        subprocess.run([exe, out], check=True)
        print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
