"""CPU checks of the drop-in boundary: libkml.so builds for sm_100a without a
GPU, exports every symbol include/kml.h declares, keeps struct layouts in sync
with the Python mirror, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    import kml
    kml.build()
    lib = kml.lib()
    hdr = open(os.path.join(ROOT, "include", "kml.h")).read()
    declared = set(re.findall(r"\b(kml_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(kml.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    import kml
    assert C.sizeof(kml.Result) == 4 * 8 + 8 + 4 * 4 + 9 * 8 + 12 * 8
    p = kml.default_params()
    assert p.max_db_results == 50 and p.lowe_ratio == 0.9 and p.ransac_seed == 12345
    assert p.top_k_verify == 16 and abs(p.ransac_probability_mono - 0.995) < 1e-15
    assert C.sizeof(kml.Params) == 168 and p.matcher_norm == 0 and p.matcher_engine == 1 and p.mono_algorithm == 0


def test_no_cpu_fallback_without_device():
    import kml
    if kml.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(kml.KmlError) as e:
        kml.LoopClosureDetector()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "kimera-multi_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "kml_oracle" not in src and "kmo_" not in src and "oracle/" not in src, (dirpath, f)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libkml.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out


def run_c_client(tmp_path):
    """Compile tests/abi_client.c as strict C99 against include/kml.h, link it to libkml.so and run it."""
    import subprocess
    import kml
    kml.build()
    pkg = os.path.join(ROOT, "kimera-multi_b200")
    exe = str(tmp_path / "abi_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_client.c"), "-o", exe, "-L", pkg, "-lkml",
                    "-Wl,-rpath," + pkg], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/kml.h under gcc -std=c99 -pedantic -Werror (and as C++11); a C program with no Python
    or torch in the process links libkml.so and, on a box without a GPU, is refused loudly."""
    import subprocess
    import kml
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++",
                    os.path.join(ROOT, "include", "kml.h")], check=True)
    out = run_c_client(tmp_path)
    if kml.device_count() == 0:
        assert "no device, create refused" in out and "no CPU fallback" in out
