"""The whole library on the CPU: tests/emu/build_libkml_emu.py compiles EVERY source of libkml
(C ABI, host orchestration, kernels) for the host — kernels under the SIMT emulator
(tests/emu/cuda_emu.h), the CUDA runtime replaced by tests/emu/fake_cudart.cpp — into
libkml_emu.so, which exports the product's C ABI.  The `-m gpu` parity tests then run UNCHANGED
against it in a subprocess (pytest plugin tests/emu/emu_plugin.py points the ctypes mirror at the
emulated build), and so does the plain-C client.  This is how the code written after the last GPU
minute of a round is exercised before the next GPU run: it checks logic (host code, kernel
semantics, the C ABI end to end), not the hardware — the GPU run of the same tests stays the
parity gate.  The product never loads this library: kml/_lib.py has no switch for it."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")

# the GPU tests that finish in seconds under emulation (the full-size ones of test_gpu_fullsize.py
# would take many minutes on one host core)
EMULATED_GPU_TESTS = [
    "tests/test_golden.py",
    "tests/test_gpu_parity.py",
    "tests/test_postfilter.py",
    "tests/test_vocab.py",
    "tests/test_wide_database.py",
    "tests/test_zy_edge_cases.py",
    "tests/test_gpu_sharded.py",
]


@pytest.fixture(scope="module")
def libkml_emu(tmp_path_factory):
    sys.path.insert(0, EMU)
    try:
        import build_libkml_emu
        return build_libkml_emu.build(str(tmp_path_factory.mktemp("libkml_emu")))
    finally:
        sys.path.remove(EMU)


def test_emulated_library_exports_the_c_abi(libkml_emu):
    import ctypes as C
    import kml
    lib = C.CDLL(libkml_emu)
    for name in kml.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.kml_device_count() == 1


def test_plain_c_client_against_the_emulated_library(libkml_emu, tmp_path):
    """tests/abi_client.c, GPU branch: frames in, CUDA matcher (emulated), matched indices out."""
    exe = str(tmp_path / "abi_client_emu")
    d = os.path.dirname(libkml_emu)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_client.c"), "-o", exe, "-L", d, "-lkml_emu", "-Wl,-rpath," + d],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "64 matches through the CUDA matcher" in r.stdout, r.stdout + r.stderr


def test_gpu_parity_tests_against_the_emulated_library(libkml_emu):
    """`pytest -m gpu` of the parity tests, unchanged, with libkml_emu.so behind the ctypes mirror."""
    env = dict(os.environ, KML_EMU_LIB=libkml_emu,
               PYTHONPATH=EMU + os.pathsep + os.environ.get("PYTHONPATH", ""))
    # not test_query_lanes_concurrent: the emulator serialises launches, so concurrent lanes show nothing
    # new there and take minutes; four xdist workers, each a process with its own emulator
    r = subprocess.run([sys.executable, "-m", "pytest", *EMULATED_GPU_TESTS, "-m", "gpu", "-p", "emu_plugin",
                        "-q", "-p", "no:cacheprovider", "-n", "4", "--timeout", "600", "--timeout-method", "thread",
                        "--deselect", "tests/test_gpu_parity.py::test_query_lanes_concurrent",
                        # CUDA graphs: nothing to replay on the emulator, and thirteen 24-query batches take minutes
                        "--deselect", "tests/test_gpu_parity.py::test_throughput_batch_graph_replay",
                        # needs two real GPUs and NCCL; its device half (merge kernel) runs here
                        "--deselect", "tests/test_gpu_sharded.py::test_two_gpu_sharded_query_equals_merge_of_local_records"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 24, tail
    assert "failed" not in r.stdout.splitlines()[-1] and "skipped" not in r.stdout.splitlines()[-1], tail


def test_smoke_entry_point_against_the_emulated_library(libkml_emu):
    """__graft_entry__.smoke() (what the driver runs on cuda:0 before the bench) with the emulated
    build behind the mirror: database fill, a 4-query batch, record-by-record parity with the oracle."""
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); import __graft_entry__ as g; import kml._lib as L; "
            "lib = C.CDLL(%r); lib.kml_last_error.restype = C.c_char_p; lib.kml_last_error.argtypes = [C.c_void_p]; "
            "L._lib = lib; g.smoke()" % (ROOT, libkml_emu))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout and "parity with oracle" in r.stdout, r.stdout + r.stderr
