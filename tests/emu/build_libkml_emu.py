"""Builds libkml_emu.so: EVERY source of libkml (kimera-multi_b200/csrc/*.cu — C ABI, host
orchestration and kernels) compiled for the host with g++, the kernels running under
tests/emu/cuda_emu.h and the CUDA runtime replaced by tests/emu/fake_cudart.cpp.  The result
exports the same C ABI as libkml.so and is test infrastructure only: it is never installed next
to the product and nothing in kimera-multi_b200/ knows about it.

    python tests/emu/build_libkml_emu.py OUT_DIR  ->  OUT_DIR/libkml_emu.so
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "kimera-multi_b200", "csrc")
SOURCES = ["api_core", "hamming", "hamming_tc", "comm", "bow", "ransac", "lcd", "vocab", "select", "postfilter", "persist"]  # = the Makefile's OBJS
# -ffp-contract=off everywhere: what -fmad=false is for ransac.cu, and the other sources hold no
# multiply-add that the GPU build may contract into an FMA with a different result
FLAGS = ["-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
         "-Wno-unused-variable", "-Wno-attributes", "-I/usr/local/cuda/include", "-I", HERE,
         "-I", os.path.join(ROOT, "include")]


def build(out_dir):
    os.makedirs(out_dir, exist_ok=True)

    def cc(name):
        obj = os.path.join(out_dir, name + ".o")
        src = os.path.join(CSRC, name + ".cu") if name != "fake_cudart" else os.path.join(HERE, "fake_cudart.cpp")
        pre = ["-x", "c++", "-include", os.path.join(HERE, "cuda_emu.h")] if name != "fake_cudart" else []
        subprocess.run(["g++"] + FLAGS + pre + ["-c", src, "-o", obj], check=True)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, SOURCES + ["fake_cudart"]))
    so = os.path.join(out_dir, "libkml_emu.so")
    # -Bsymbolic: the library's calls to cuda* bind to fake_cudart.o even if a real libcudart is loaded in the process
    subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", so] + objs + ["-ldl"], check=True)
    return so


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/tmp/kml_emu"))
