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


def test_params_from_the_reference_yaml(tmp_path):
    """kml_params_from_yaml on the reference's own parameter file (OpenCV FileStorage dialect) and on
    a copy of its text when the reference tree is not there (GPU box)."""
    import kml
    ref = "/root/reference/params/D455/LcdParams.yaml"
    text = ("%YAML:1.0\n# LoopClosureDetector parameters:\nuse_nss: 1\nalpha: 0.4\nmin_temporal_matches: 1\n"
            "recent_frames_window: 100\nmax_db_results: 50\nmin_nss_factor: 0.05\nmin_matches_per_island: 1\n"
            "max_intraisland_gap: 3\nmax_nrFrames_between_islands: 3\nmax_nrFrames_between_queries: 2\n\n"
            "lowe_ratio: 0.7\nmatcher_type: 3\nnfeatures: 700\nmin_nr_2d2d_inliers: 10\nmin_nr_3d3d_inliers: 5\n"
            "ransac_threshold_2d2d: 1e-06\nransac_threshold_3d3d: 0.3\nransac_use_1point_3d3d: 1  # if 1, use rotation\n"
            "ransac_max_iterations: 500\nransac_probability: 0.995\nransac_randomize: 0\n"
            "ransac_2d2d_algorithm: 0 # Stewenius\n")
    paths = [str(tmp_path / "LcdParams.yaml")]
    open(paths[0], "w").write(text)
    if os.path.exists(ref):
        paths.append(ref)
    for path in paths:
        p = kml.params_from_yaml(path)
        assert p.alpha == 0.4 and p.max_db_results == 50 and p.min_nss_factor == 0.05 and p.dist_local == 100
        assert p.lowe_ratio == 0.7 and p.ransac_threshold_mono == 1e-6 and p.ransac_threshold == 0.3
        assert p.max_ransac_iterations == 500 and p.max_ransac_iterations_mono == 500 and p.ransac_probability == 0.995
        assert p.geometric_verification_min_inlier_count == 5 and p.ransac_randomize == 0
        assert p.ransac_use_1point_3d3d == 1 and p.mono_algorithm == 1           # 0 = STEWENIUS in OpenGV's enum
        assert p.matcher_norm == 0 and p.n_mapped >= 19                          # "3: BRUTEFORCE_HAMMING" per the file's comment
        q = kml.params_from_yaml(path, literal_matcher_enum=True)                # create(3) = BRUTEFORCE_L1
        assert q.matcher_norm == 1 and q.matcher_engine == 0
    with pytest.raises(kml.KmlError):
        kml.params_from_yaml(str(tmp_path / "missing.yaml"))
