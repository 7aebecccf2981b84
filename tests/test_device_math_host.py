"""The device source of the RANSAC arithmetic (kimera-multi_b200/csrc/geom.cuh and
fivept_thread.cuh — the text nvcc compiles into libkml.so) compiled for the HOST by
tests/device_math_host.cpp, one thread at a time, and held to the oracle bit for bit.  Runs
without a GPU: a change to the device arithmetic that breaks the contract of DESIGN.md §4 fails
here, before any GPU time is spent.  (The GPU tests check the same through the kernels.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THR = 1e-6


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def dmh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dmh") / "libdmh.so")
    # -ffp-contract=off is the host equivalent of nvcc -fmad=false (kimera-multi_b200/Makefile)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-shared", "-fPIC", "-Wall",
                    "-Wno-unknown-pragmas", "-Wno-unused-function",
                    os.path.join(ROOT, "tests", "device_math_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.dmh_mono_residual.restype = C.c_double
    lib.dmh_arun_sqdist.restype = C.c_double
    return lib


def _two_view(rng, N, kind="plain"):
    from test_gpu_parity import _nister_case
    a, b = _nister_case(rng, N, kind)
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def _device_mono_model(dmh, f1, f2, s, force_generic):
    m, info = np.zeros(12), np.zeros(5, np.int32)
    ok = dmh.dmh_mono_model(_dp(f1), _dp(f2), s.ctypes.data_as(C.POINTER(C.c_uint16)), int(force_generic),
                            _dp(m), info.ctypes.data_as(C.POINTER(C.c_int)))
    return bool(ok), m, info


def test_contract_rsqrt_same_bits_and_within_three_ulp(oracle, dmh):
    """krsqrt is the one non-IEEE primitive of the contract: the device source and the oracle
    return the same bits, within 3 ulp of 1/sqrt(x), over 25 decades; zero, infinity and NaN included in the
    bit comparison."""
    rng = np.random.default_rng(4)
    x = rng.uniform(1.0, 2.0, 200000) * 2.0 ** rng.integers(-400, 400, 200000)
    x = np.concatenate([x, [0.0, 1.0, 4.0, 1e-300, 1e300, np.inf, np.nan]])
    a = oracle.krsqrt(x)
    b = np.zeros_like(x)
    dmh.dmh_krsqrt(_dp(x), C.c_int(x.size), _dp(b))
    assert a.tobytes() == b.tobytes()
    fin = np.isfinite(x) & (x > 0)
    ref = 1.0 / np.sqrt(x[fin].astype(np.longdouble))
    err = np.abs((a[fin] - ref) / ref).astype(np.float64)
    assert err.max() <= 3 * 2.0 ** -53, err.max()
    assert a[x == 1.0][0] > 0.999999999999999 and abs(a[x == 4.0][0] - 0.5) < 1e-15


def test_svd3_both_device_variants_equal_the_oracle(oracle, dmh):
    rng = np.random.default_rng(5)
    for k in range(3000):
        A = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-3, 4)
        if k % 5 == 0:
            A[:, 2] = A[:, 0] * rng.normal()          # rank 2
        if k % 17 == 0:
            A[:, 1] = A[:, 0]; A[:, 2] = -A[:, 0]     # rank 1
        if k == 1:
            A[:] = 0.0
        if k % 7 == 0:                                # an essential matrix: two equal singular values
            A = np.cross(np.eye(3), rng.normal(size=3)) @ Rot.from_rotvec(rng.normal(size=3)).as_matrix()
        A = np.ascontiguousarray(A)
        Uo, So, Vo = oracle.svd3(A)
        for fn in (dmh.dmh_svd3, dmh.dmh_svd3_r):
            U, S, V = np.zeros(9), np.zeros(3), np.zeros(9)
            fn(_dp(A), _dp(U), _dp(S), _dp(V))
            assert np.array_equal(U, Uo.ravel()) and np.array_equal(S, So) and np.array_equal(V, Vo.ravel()), k


def test_arun_model_and_residual_equal_the_oracle(oracle, dmh):
    rng = np.random.default_rng(6)
    for k in range(2000):
        p2 = rng.uniform(-5, 5, (3, 3))
        R = Rot.from_rotvec(rng.normal(size=3) * 0.3).as_matrix()
        p1 = p2 @ R.T + rng.uniform(-1, 1, 3) + rng.normal(size=(3, 3)) * 0.02
        if k % 10 == 0:
            p1[2], p2[2] = p1[1], p2[1]               # repeated point: rank-deficient H
        if k % 25 == 0:
            p1[:], p2[:] = 0.0, 0.0                   # invalid depth on all three
        p1, p2 = np.ascontiguousarray(p1), np.ascontiguousarray(p2)
        M = np.zeros(12)
        dmh.dmh_arun3(_dp(p1), _dp(p2), _dp(M))
        Mo = oracle.arun3(p1, p2).ravel()
        assert np.array_equal(M, Mo, equal_nan=True), k
        if np.isfinite(Mo).all():
            a, b = rng.uniform(-5, 5, 3), rng.uniform(-5, 5, 3)
            d2 = dmh.dmh_arun_sqdist(_dp(M), _dp(a), _dp(b))
            assert np.sqrt(d2) == oracle.arun_residual(Mo, a, b)


def test_mono_draw_pipeline_equals_the_oracle(oracle, dmh):
    """front -> isolate (+ deferred bisections) -> items -> winner, i.e. what the four mono kernels
    do for one draw, against kmo_mono_model: validity and all 12 model entries bit-identical, on
    the generic fast path and with the generic variable-degree path forced, on plain and awkward
    scenes (every `kind` of the GPU corner-case test) and on samples with repeated indices."""
    rng = np.random.default_rng(7)
    n_models = n_deferred = n_valid = 0
    for kind in ("plain", "low_parallax", "far_points", "near_centres", "non_unit", "duplicates"):
        for trial in range(40):
            f1, f2 = _two_view(rng, 60, kind)
            for d in range(10):
                s = rng.permutation(60)[:8].astype(np.uint16)
                if d == 9:
                    s[1] = s[0]                       # degenerate sample
                ok0, m0 = oracle.mono_model(f1, f2, s)
                for fg in (False, True):
                    ok1, m1, info = _device_mono_model(dmh, f1, f2, s, fg)
                    tag = (kind, trial, d, fg)
                    assert ok1 == ok0, tag
                    if ok0:
                        assert np.array_equal(m1, m0.ravel()), tag
                    n_models += 1
                    n_valid += ok0
                    n_deferred += int(info[2] != 0)
    assert n_models == 4800 and n_valid > 4000 and n_deferred > 100   # both isolation paths were exercised


def test_mono_residual_and_fast_inlier_filter(oracle, dmh):
    """mono_residual bit-identical to the oracle's; the approximate inlier filter never contradicts
    the exact test residual < threshold (it may only answer 'undecided'), and it does decide the
    bulk of the correspondences on well-conditioned scenes."""
    rng = np.random.default_rng(8)
    decided = total = 0
    for kind in ("plain", "low_parallax", "far_points", "near_centres", "duplicates"):
        for trial in range(20):
            f1, f2 = _two_view(rng, 80, kind)
            s = rng.permutation(80)[:8].astype(np.uint16)
            ok, M = oracle.mono_model(f1, f2, s)
            if not ok:
                continue
            M = np.ascontiguousarray(M.ravel())
            for thr in (1e-6, 1e-9, 1e-4, 5e-8):
                for i in range(80):
                    r0 = oracle.mono_residual(M, f1[i], f2[i])
                    r1 = dmh.dmh_mono_residual(_dp(M), _dp(f1[i]), _dp(f2[i]))
                    assert r0 == r1 or (np.isnan(r0) and np.isnan(r1)), (kind, trial, i)
                    fast = dmh.dmh_mono_inlier_fast(_dp(M), _dp(f1[i]), _dp(f2[i]), C.c_double(thr))
                    assert fast in (-1, 0, 1)
                    if fast >= 0:
                        assert bool(fast) == bool(r0 < thr), (kind, trial, i, thr, r0)
                    if kind == "plain":
                        decided += fast >= 0
                        total += 1
    assert decided > 0.9 * total
