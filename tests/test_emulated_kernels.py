"""CUDA kernels of libkml run on the CPU: tests/emu/cuda_emu.h executes a __global__ function of
this repo — the source nvcc compiles, not a restatement — one CTA at a time with every CUDA thread
as a cooperative fibre (barriers, full-mask shuffles / ballots, shared-memory atomics), behind
the product's own host code for that kernel.  No GPU needed: this is where a change to a kernel's
logic fails first.  It models semantics, not timing or the memory system; the GPU tests remain
the parity tests proper.

Covered here: bow_score_kernel + the CSR build, tiling and tile merge of csrc/bow_merge.h,
including databases wider than one accumulator tile (BASELINE.json configs[4], 50 000 keyframes
per database), which no GPU run of round 1 reached."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOW_RTOL = 1e-6


def _P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _build_harness(tmp_path_factory, name, extra=()):
    """tests/emu/<name>.cpp (which includes the .cu under test) + the fake CUDA runtime -> a shared library"""
    so = str(tmp_path_factory.mktemp("emu") / ("lib%s.so" % name))
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas",
                    "-Wno-unused-function", "-Wno-unused-variable", "-Wno-attributes", *extra,
                    "-I/usr/local/cuda/include", "-I", emu, os.path.join(emu, name + ".cpp"),
                    os.path.join(emu, "fake_cudart.cpp"), "-o", so], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module")
def bowemu(tmp_path_factory):
    lib = _build_harness(tmp_path_factory, "bow_emu")
    lib.bowemu_db_create.restype = C.c_void_p
    return lib


class EmuDb:
    def __init__(self, lib):
        self.lib, self.h = lib, C.c_void_p(lib.bowemu_db_create())

    def add_bulk(self, off, ids, vals):
        off, ids = np.ascontiguousarray(off, np.int64), np.ascontiguousarray(ids, np.uint32)
        vals = np.ascontiguousarray(vals, np.float32)
        self.lib.bowemu_db_add(self.h, len(off) - 1, _P(off, C.c_int64), _P(ids, C.c_uint32), _P(vals, C.c_float))

    def close(self):
        self.lib.bowemu_db_destroy(self.h)


def _csr(vecs):
    off = np.zeros(len(vecs) + 1, np.int64)
    off[1:] = np.cumsum([len(v[0]) for v in vecs])
    ids = np.concatenate([v[0] for v in vecs] + [np.zeros(0, np.uint32)]).astype(np.uint32)
    vals = np.concatenate([v[1] for v in vecs] + [np.zeros(0, np.float32)]).astype(np.float32)
    return off, ids, vals


def emu_query(lib, dbs, queries, K, max_id=None, prev=None, tile_cap=0):
    """run_bow (lcd.cu) with the kernel emulated: returns entries [B][n_db][K], scores, counts, nss, tiles."""
    B, n_db = len(queries), len(dbs)
    off, ids, vals = _csr(queries)
    arr = (C.c_void_p * n_db)(*[d.h for d in dbs])
    oe, osc = np.zeros((B, n_db, K), np.uint32), np.zeros((B, n_db, K))
    oc, nss = np.zeros((B, n_db), np.int32), np.zeros(B)
    mid = None if max_id is None else np.ascontiguousarray(max_id, np.int32)
    pa = (None, None, None)
    if prev is not None:
        poff, pids, pvals = _csr(prev)
        pa = (_P(poff, C.c_int64), _P(pids, C.c_uint32), _P(pvals, C.c_float))
    touched = C.c_ulonglong(0)
    nt = lib.bowemu_query(arr, n_db, B, _P(off, C.c_int64), _P(ids, C.c_uint32), _P(vals, C.c_float), *pa, K,
                          None if mid is None else _P(mid, C.c_int32), tile_cap, _P(oe, C.c_uint32),
                          _P(osc, C.c_double), _P(oc, C.c_int32), _P(nss, C.c_double), C.byref(touched))
    return oe, osc, oc, nss, nt, touched.value


def _check_against_oracle(db, qi, qv, K, max_id, e1, s1):
    from test_gpu_parity import _same_order_modulo_ties
    e0, s0 = db.query(qi, qv, K, max_id)
    assert len(e0) == len(e1), (len(e0), len(e1))
    np.testing.assert_allclose(s1, s0, rtol=BOW_RTOL, atol=0)
    assert _same_order_modulo_ties(e0, s0, e1, s1)
    return len(e0)


def _bow_tile(n, cap=20480):
    """bow_tiling of csrc/bow_merge.h: equal tiles of at most `cap` entries, a multiple of 256"""
    nt = (n + cap - 1) // cap
    t = (n + nt - 1) // nt
    return max(256, (t + 255) // 256 * 256)


def test_bow_kernel_database_spanning_several_entry_tiles(oracle, bowemu):
    """The scenario of tests/test_wide_database.py (one database of 60 000 entries = three
    equal entry tiles, exact duplicates on both sides of every tile boundary, max_id cuts inside
    the first and the last tile) through the emulated kernel and the product's tile merge."""
    rng = np.random.default_rng(77)
    n, words, vocab = 60000, 40, 3000
    base = rng.integers(0, vocab // words, (n, words))
    ids = (np.arange(words)[None, :] * (vocab // words) + base).astype(np.uint32)
    vals = rng.random((n, words)).astype(np.float32) + np.float32(0.01)
    vals = (vals / vals.sum(axis=1, keepdims=True)).astype(np.float32)
    T = _bow_tile(n)
    dup = [5, T - 1, T, 2 * T - 1, 2 * T, 59999]
    for e in dup[1:]:
        ids[e], vals[e] = ids[dup[0]], vals[dup[0]]
    det = EmuDb(bowemu)
    det.add_bulk(np.arange(n + 1) * words, ids.reshape(-1), vals.reshape(-1))
    db = oracle.Database()
    for i in range(n):
        db.add(ids[i], vals[i])
    queries = [(ids[dup[0]], vals[dup[0]])]
    for k in range(3):
        qv = rng.random(words).astype(np.float32)
        queries.append((ids[1000 + 19000 * k].copy(), (qv / qv.sum()).astype(np.float32)))
    for K, max_id in [(50, -1), (128, -1), (1, -1), (50, 30000), (50, 100), (50, 59000)]:
        oe, osc, oc, _, nt, touched = emu_query(bowemu, [det], queries, K, [max_id])
        assert nt == 3
        assert touched == sum(int(np.isin(ids, q[0]).sum()) for q in queries)   # algorithmic postings, counted once
        for b, (qi, qv) in enumerate(queries):
            c = oc[b, 0]
            assert _check_against_oracle(db, qi, qv, K, max_id, oe[b, 0, :c], osc[b, 0, :c]) > 0
    oe, osc, oc, _, _, _ = emu_query(bowemu, [det], queries[:1], 50, [-1])
    assert list(oe[0, 0, :len(dup)]) == dup and np.all(np.abs(osc[0, 0, :len(dup)] - 1.0) < 1e-6)
    det.close()


def test_bow_kernel_ragged_databases_small_tiles_and_nss(oracle, bowemu):
    """Several databases of different sizes in one launch (empty, smaller than a tile, many tiles —
    the tile is shrunk to 256 entries), per-database max_id, the NSS factor of the batch, empty
    and 1 024-word query vectors; single-tile launch of the same data for comparison."""
    rng = np.random.default_rng(3)
    vocab = 5000

    def vec(nw):
        i = np.sort(rng.choice(vocab, nw, replace=False)).astype(np.uint32)
        v = rng.random(nw).astype(np.float32) + np.float32(1e-3)
        return i, (v / v.sum()).astype(np.float32)

    sizes = [0, 100, 700, 1500]
    dbs, refs = [], []
    for n in sizes:
        d, r = EmuDb(bowemu), oracle.Database()
        vecs = [vec(int(rng.integers(1, 60))) for _ in range(n)]
        if n:
            d.add_bulk(*_csr(vecs))
        for v in vecs:
            r.add(*v)
        dbs.append(d)
        refs.append(r)
    queries = [vec(50), vec(1024), (np.zeros(0, np.uint32), np.zeros(0, np.float32)), vec(1), vec(300)]
    prev = [vec(40), queries[1], vec(10), (np.zeros(0, np.uint32), np.zeros(0, np.float32)), vec(300)]
    prev[4] = (queries[4][0].copy(), prev[4][1])                      # same words, other weights
    max_id = [-1, 37, 650, -1]
    for K in (50, 7):
        outs = []
        for tile_cap in (256, 0):
            oe, osc, oc, nss, nt, _ = emu_query(bowemu, dbs, queries, K, max_id, prev, tile_cap)
            assert nt == (6 if tile_cap else 1)
            outs.append((oe.copy(), osc.copy(), oc.copy(), nss.copy()))
            for b, (qi, qv) in enumerate(queries):
                assert nss[b] == pytest.approx(oracle.bow_score(qi, qv, *prev[b]), rel=BOW_RTOL, abs=0)
                for d in range(len(dbs)):
                    c = oc[b, d]
                    if sizes[d] == 0 or len(qi) == 0:
                        assert c == 0
                        continue
                    _check_against_oracle(refs[d], qi, qv, K, max_id[d], oe[b, d, :c], osc[b, d, :c])
        # fixed-point accumulation: the tiled and the single-tile launch agree bit for bit
        for x, y in zip(outs[0], outs[1]):
            assert np.array_equal(x, y)
    assert outs[0][3][1] == pytest.approx(1.0, abs=1e-6)              # a vector against itself
    for d in dbs:
        d.close()


# ---------------------------------------------------------------------------------- matcher
@pytest.fixture(scope="module")
def hamemu(tmp_path_factory):
    return _build_harness(tmp_path_factory, "hamming_emu")


def emu_knn2(lib, q, t, norm=0, range_len=1 << 20):
    q, t = np.ascontiguousarray(q, np.uint8).reshape(-1, 32), np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx, dist = np.zeros((len(q), 2), np.uint32), np.zeros((len(q), 2), np.uint16)
    lib.hamemu_knn2(_P(q, C.c_uint8), len(q), _P(t, C.c_uint8), len(t), norm, range_len,
                    _P(idx, C.c_uint32), _P(dist, C.c_uint16))
    return idx, dist


def emu_match_lowe(lib, q, t, lowe, norm=0):
    """q: [P][nq][32], t: [P][nt][32] -> list of (iq, im) per pair"""
    q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
    Pn, nq, nt = q.shape[0], q.shape[1], t.shape[1]
    iq, im = np.zeros((Pn, max(nq, 1)), np.uint16), np.zeros((Pn, max(nq, 1)), np.uint16)
    M = np.zeros(Pn, np.int32)
    lib.hamemu_match_lowe(_P(q, C.c_uint8), nq, _P(t, C.c_uint8), nt, Pn, norm, C.c_double(lowe),
                          _P(iq, C.c_uint16), _P(im, C.c_uint16), _P(M, C.c_int32))
    return [(iq[p, :M[p]].astype(np.uint32), im[p, :M[p]].astype(np.uint32)) for p in range(Pn)]


def test_matcher_kernels_against_the_bfmatcher_golden_vectors(hamemu):
    """hamming_knn2_kernel (Hamming and byte-wise L1) + knn2_reduce_kernel + lowe_compact_kernel,
    emulated, against the fixture OpenCV's BFMatcher wrote (tests/golden): the same vectors the
    GPU test holds the compiled kernels to."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    for range_len in (1 << 20, 64):                 # one CTA, and many ranges merged by the reduce kernel
        idx, dist = emu_knn2(hamemu, G["knn_q"], G["knn_t"], 0, range_len)
        assert np.array_equal(idx, G["ham_idx"]) and np.array_equal(dist, G["ham_dist"])
        idx, dist = emu_knn2(hamemu, G["knn_q"], G["knn_t"], 1, range_len)
        assert np.array_equal(idx, G["l1_idx"]) and np.array_equal(dist, G["l1_dist"])
    (iq, im), = emu_match_lowe(hamemu, G["knn_q"][None], G["knn_t"][None], 0.8)
    assert np.array_equal(np.c_[iq, im], G["lowe08_pairs"])


def test_matcher_kernels_edge_cases_against_the_oracle(oracle, hamemu):
    """Sizes around the kernel's tiling (512 queries per pass, 512-descriptor TMA stages, the 4-way
    unrolled inner loop), duplicates and exact hits (ties -> lowest train index), k > nTrain, an
    empty train set, and the Lowe compaction over several 256-query passes for several pairs."""
    rng = np.random.default_rng(12)
    for nq, nt in [(1, 1), (3, 2), (5, 7), (500, 500), (513, 511), (37, 1030), (1025, 3), (4, 0), (0, 5)]:
        q = rng.integers(0, 256, (nq, 32), np.uint8)
        t = rng.integers(0, 256, (nt, 32), np.uint8)
        if nt >= 6 and nq:
            t[4] = t[1]; t[5] = t[1]; q[0] = t[1]
        for norm, ref in ((0, oracle.hamming_knn2), (1, oracle.l1_knn2)):
            if nq == 0:
                emu_knn2(hamemu, q, t, norm)          # nothing to do, must not hang
                continue
            i0, d0 = ref(q, t)
            for range_len in (1 << 20, 200):
                i1, d1 = emu_knn2(hamemu, q, t, norm, range_len)
                assert np.array_equal(i0, i1) and np.array_equal(d0, d1), (nq, nt, norm, range_len)
    # Lowe + compaction: related frames so that a good share of the queries survive
    Pn, nq, nt = 3, 700, 650
    t = rng.integers(0, 256, (Pn, nt, 32), np.uint8)
    q = rng.integers(0, 256, (Pn, nq, 32), np.uint8)
    for p in range(Pn):
        sel = rng.permutation(nq)[:400]
        src = rng.integers(0, nt, 400)
        q[p, sel] = t[p, src] ^ (rng.random((400, 32)) < 0.03).astype(np.uint8)   # a few flipped bits
    for norm in (0, 1):
        for lowe in (0.9, 0.5):
            got = emu_match_lowe(hamemu, q, t, lowe, norm)
            for p in range(Pn):
                if norm == 0:
                    iq0, im0 = oracle.match_lowe(q[p], t[p], lowe)
                else:                                   # the oracle's Lowe helper is Hamming: restate it on l1_knn2
                    i0, d0 = oracle.l1_knn2(q[p], t[p])
                    keep = d0[:, 0].astype(np.float64) < lowe * d0[:, 1].astype(np.float64)
                    iq0, im0 = np.nonzero(keep)[0], i0[keep, 0]
                assert len(iq0) > 100
                assert np.array_equal(got[p][0], iq0) and np.array_equal(got[p][1], im0), (norm, lowe, p)


# ----------------------------------------------------------------------------------- RANSAC
@pytest.fixture(scope="module")
def ransacemu(tmp_path_factory):
    # -ffp-contract=off: the host equivalent of the -fmad=false ransac.cu is compiled with
    return _build_harness(tmp_path_factory, "ransac_emu", extra=("-ffp-contract=off",))


def emu_ransac(lib, mono, a, b, thr, prob=0.995, max_it=1000, seed=12345, full=False, force_generic=False):
    """kml_ransac_nister_batch / kml_ransac_arun_batch with every kernel emulated; a, b: [P][N][3]."""
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    Pn, N = a.shape[0], a.shape[1]
    models = np.zeros((Pn, 12))
    ninl, it, bd = np.zeros(Pn, np.int32), np.zeros(Pn, np.int32), np.zeros(Pn, np.int32)
    mask = np.zeros((Pn, max((N + 31) // 32, 1)), np.uint32)
    rc = lib.ransacemu_batch(int(mono), Pn, N, _P(a, C.c_double), _P(b, C.c_double), C.c_double(thr),
                             C.c_double(prob), max_it, C.c_uint32(seed), int(full), int(force_generic),
                             _P(models, C.c_double), _P(ninl, C.c_int32), _P(it, C.c_int32), _P(bd, C.c_int32),
                             _P(mask, C.c_uint32))
    assert rc == 0
    return dict(models=models, n_inliers=ninl, iterations=it, best_draw=bd, mask=mask)


def _mask_to_indices(m, N):
    return np.nonzero(np.unpackbits(m.view(np.uint8), bitorder="little")[:N])[0]


def _same_as_oracle(o, g, p, N, tag):
    assert o["iterations"] == g["iterations"][p], tag
    assert o["best_draw"] == g["best_draw"][p], tag
    assert o["n_inliers"] == g["n_inliers"][p], tag
    assert np.array_equal(o["inliers"], _mask_to_indices(g["mask"][p], N)), tag
    if o["best_draw"] >= 0:
        assert np.array_equal(o["model"].ravel(), g["models"][p]), tag      # all 12 entries, bit for bit


def _scene(rng, N, outlier_rate):
    from scipy.spatial.transform import Rotation as Rot
    X = np.c_[rng.uniform(-5, 5, N), rng.uniform(-5, 5, N), rng.uniform(2, 12, N)]
    R = Rot.from_rotvec(rng.normal(size=3) * 0.2).as_matrix()
    t = rng.uniform(-1, 1, 3)
    X2 = (X - t) @ R + rng.normal(size=X.shape) * 0.02
    out = rng.random(N) < outlier_rate
    X2[out] = rng.uniform(-8, 8, (int(out.sum()), 3))
    f1 = X / np.linalg.norm(X, axis=1, keepdims=True)
    f2 = X2 + rng.normal(size=X.shape) * 1e-3
    f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
    return X, X2, f1, f2


def test_stereo_ransac_kernels_equal_the_oracle_loop(oracle, ransacemu):
    """sac_init -> 6 x (stereo_chunk + sac_replay) -> sac_select, emulated, against the oracle's
    sequential Ransac::computeModel: iteration counts, winning draw, inlier set and model bit-exact,
    from clean to outlier-heavy problems (early exit to all 1 001 trials), invalid depths, fewer
    points than a sample, and the full-hypotheses mode of BASELINE.json configs[3]."""
    rng = np.random.default_rng(21)
    Pn, N = 6, 120
    p1, p2 = np.zeros((Pn, N, 3)), np.zeros((Pn, N, 3))
    for p in range(Pn):
        p1[p], p2[p], _, _ = _scene(rng, N, [0.05, 0.2, 0.4, 0.6, 0.8, 0.97][p])
        if p == 2:
            p1[p, ::7] = 0.0                                     # invalid depth on the query side
    for thr in (0.5, 0.05):
        g = emu_ransac(ransacemu, False, p1, p2, thr)
        for p in range(Pn):
            _same_as_oracle(oracle.ransac_arun(p1[p], p2[p], thr, 0.995, 1000, 12345), g, p, N, (thr, p))
    assert g["iterations"].min() < 100 and g["iterations"].max() == 1001
    g = emu_ransac(ransacemu, False, p1, p2, 0.5, seed=7, max_it=300, prob=0.99)
    for p in range(Pn):
        _same_as_oracle(oracle.ransac_arun(p1[p], p2[p], 0.5, 0.99, 300, 7), g, p, N, ("seed 7", p))
    g = emu_ransac(ransacemu, False, p1, p2, 0.5, full=True)
    assert (g["iterations"] == 1001).all()
    g = emu_ransac(ransacemu, False, p1[:, :2], p2[:, :2], 0.5)
    assert (g["best_draw"] == -1).all() and (g["n_inliers"] == 0).all()


def test_mono_ransac_kernels_equal_the_oracle_loop(oracle, ransacemu):
    """sac_init -> 6 x (mono_front, mono_isolate, mono_isolate_deferred, mono_item, mono_count,
    sac_replay) -> sac_select, emulated, against the oracle's loop: the same bit-exact outcome on
    the awkward scenes of the GPU corner-case test (every guard of the fast inlier filter), at
    several thresholds, with the generic root-isolation path forced, and with too few points."""
    from test_gpu_parity import _nister_case
    rng = np.random.default_rng(2024)
    kinds = ["plain", "low_parallax", "far_points", "near_centres", "non_unit", "duplicates"]
    # fg: 1 = generic stage-2 path; 2 = no 256-cell grid on either side, so that every chain the first grid
    # does not separate goes through the warp-cooperative Sturm bisection (rare otherwise)
    for thr, N, fg in [(1e-6, 60, 0), (1e-9, 33, 0), (1e-4, 9, 0), (5e-8, 8, 0), (1e-6, 40, 1), (1e-6, 48, 2), (1e-7, 30, 3)]:
        f1, f2 = np.zeros((len(kinds), N, 3)), np.zeros((len(kinds), N, 3))
        for i, kind in enumerate(kinds):
            f1[i], f2[i] = _nister_case(rng, N, kind)
        g = emu_ransac(ransacemu, True, f1, f2, thr, force_generic=fg)
        oracle.debug_root_grid2(0 if fg & 2 else 1)
        try:
            for i, kind in enumerate(kinds):
                _same_as_oracle(oracle.ransac_nister(f1[i], f2[i], thr, 0.995, 1000, 12345), g, i, N, (thr, N, kind, fg))
        finally:
            oracle.debug_root_grid2(1)
    # clean scenes stop early (adaptive k), through the same replay
    Pn, N = 3, 100
    f1, f2 = np.zeros((Pn, N, 3)), np.zeros((Pn, N, 3))
    for p in range(Pn):
        _, _, f1[p], f2[p] = _scene(rng, N, 0.05 * p)
    g = emu_ransac(ransacemu, True, f1, f2, 1e-5, max_it=400)
    for p in range(Pn):
        _same_as_oracle(oracle.ransac_nister(f1[p], f2[p], 1e-5, 0.995, 400, 12345), g, p, N, ("clean", p))
    assert g["iterations"].min() < 401
    g = emu_ransac(ransacemu, True, f1[:, :5], f2[:, :5], 1e-6)
    assert (g["best_draw"] == -1).all() and (g["n_inliers"] == 0).all()


# ------------------------------------------------------------------------- schedule independence
@pytest.mark.parametrize("schedule", ["reverse", "random:3"])
def test_results_do_not_depend_on_the_thread_schedule(schedule):
    """The emulator runs a CTA's threads in ascending order between rendezvous points; a GPU runs
    them in any order.  The same tests under a descending and a per-pass random order (same
    bit-exact expectations): the fixed-point BoW accumulation, the atomically compacted work lists
    of the mono rounds and the count kernel's running bound must not care."""
    if os.environ.get("KML_EMU_SCHEDULE"):
        pytest.skip("already inside a schedule run")
    import sys
    env = dict(os.environ, KML_EMU_SCHEDULE=schedule)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-p", "no:cacheprovider",
                        "-n", "4", "-k", "not schedule"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1200)
    m = re.search(r"(\d+) passed", r.stdout)
    assert r.returncode == 0 and m and int(m.group(1)) >= 9 and "failed" not in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_bow_incremental_appends_equal_a_rebuild(oracle, bowemu):
    """BowInvFile::plan_append + bow_append_kernel (csrc/bow_merge.h, bow.cu): vectors added after
    the first query are appended to the resident inverted file in place; the scorer must return the
    oracle's results after every step, and the rows must have been appended, not rebuilt."""
    rng = np.random.default_rng(9)
    vocab, words = 400, 20

    def vec():
        ids = np.sort(rng.choice(vocab, words, replace=False)).astype(np.uint32)
        v = rng.random(words).astype(np.float32) + np.float32(0.01)
        return ids, (v / v.sum()).astype(np.float32)

    det, db = EmuDb(bowemu), oracle.Database()
    first = [vec() for _ in range(50)]
    det.add_bulk(*_csr(first))
    for i, v in first:
        db.add(i, v)
    for step in range(60):
        new = [vec() for _ in range(1 + step % 3)]
        det.add_bulk(*_csr(new))
        for i, v in new:
            db.add(i, v)
        q = vec()
        oe, osc, oc, _, _, _ = emu_query(bowemu, [det], [q], 20, [-1])
        c = oc[0, 0]
        assert _check_against_oracle(db, q[0], q[1], 20, -1, oe[0, 0, :c], osc[0, 0, :c]) > 0
    assert bowemu.bowemu_db_rebuilds(det.h) <= 2 and bowemu.bowemu_db_appends(det.h) >= 55
    det.close()


def test_bow_cut_inside_a_large_tie_group(oracle, bowemu):
    """100 copies of one vector among 300 others, K = 50: the cut falls inside a group of exactly
    equal scores larger than a warp (all eight radix passes run, then the dense sweep): the copies
    with the lowest entry ids are taken, in ascending id; K = 120 takes the whole group and fills up
    with the runners-up (early-exit path)."""
    rng = np.random.default_rng(4)
    vocab, words = 500, 16

    def vec():
        ids = np.sort(rng.choice(vocab, words, replace=False)).astype(np.uint32)
        v = rng.random(words).astype(np.float32) + np.float32(0.01)
        return ids, (v / v.sum()).astype(np.float32)

    twin = vec()
    vecs = [vec() for _ in range(300)]
    twin_at = sorted(rng.choice(400, 100, replace=False).tolist())
    allv, k = [], 0
    for e in range(400):
        if e in twin_at:
            allv.append(twin)
        else:
            allv.append(vecs[k]); k += 1
    det, db = EmuDb(bowemu), oracle.Database()
    det.add_bulk(*_csr(allv))
    for i, v in allv:
        db.add(i, v)
    oe, osc, oc, _, _, _ = emu_query(bowemu, [det], [twin], 50, [-1])
    assert oc[0, 0] == 50 and list(oe[0, 0]) == twin_at[:50] and np.all(np.abs(osc[0, 0] - 1.0) < 1e-6)
    oe, osc, oc, _, _, _ = emu_query(bowemu, [det], [twin], 120, [-1])
    assert oc[0, 0] == 120 and list(oe[0, 0, :100]) == twin_at
    assert _check_against_oracle(db, twin[0], twin[1], 120, -1, oe[0, 0, :120], osc[0, 0, :120]) == 120
    det.close()


def test_ransac_per_problem_sampler_without_the_table(oracle, ransacemu):
    """KML_NO_SAMPLE_TABLE: the warp-serial per-problem sampler of sac_init / sac_replay (what problems
    with more than 1 024 correspondences use) must give the outcome the per-N sample table gives."""
    rng = np.random.default_rng(5)
    p1, p2 = np.zeros((3, 90, 3)), np.zeros((3, 90, 3))
    for p in range(3):
        p1[p], p2[p], _, _ = _scene(rng, 90, [0.3, 0.6, 0.9][p])
    g1 = emu_ransac(ransacemu, False, p1, p2, 0.5)
    os.environ["KML_NO_SAMPLE_TABLE"] = "1"
    try:
        g0 = emu_ransac(ransacemu, False, p1, p2, 0.5)
    finally:
        del os.environ["KML_NO_SAMPLE_TABLE"]
    for k in ("iterations", "best_draw", "n_inliers", "mask", "models"):
        assert np.array_equal(g0[k], g1[k]), k
    for p in range(3):
        o = oracle.ransac_arun(p1[p], p2[p], 0.5, 0.995, 1000, 12345)
        assert o["iterations"] == g0["iterations"][p] and o["best_draw"] == g0["best_draw"][p]


def test_mono_ransac_stewenius_kernels_equal_the_oracle_loop(oracle, ransacemu):
    """Row f4, ransac_2d2d_algorithm: 0 (/root/reference/params/D455/LcdParams.yaml:73): the mono round
    with the Stewenius front stage (degree-ordered constraint matrix, full Gauss-Jordan, action matrix,
    characteristic polynomial) and eigenvector item stage, emulated, against the oracle's loop with
    its Stewenius restatement: iteration counts, winning draw, inlier set and model bit-exact."""
    from test_gpu_parity import _nister_case
    rng = np.random.default_rng(77)
    kinds = ["plain", "low_parallax", "far_points", "near_centres", "duplicates"]
    os.environ["KML_EMU_STEWENIUS"] = "1"
    try:
        for thr, N in [(1e-6, 60), (1e-8, 33), (1e-4, 9)]:
            f1, f2 = np.zeros((len(kinds), N, 3)), np.zeros((len(kinds), N, 3))
            for i, kind in enumerate(kinds):
                f1[i], f2[i] = _nister_case(rng, N, kind)
            g = emu_ransac(ransacemu, True, f1, f2, thr)
            for p in range(len(kinds)):
                o = oracle.ransac_stewenius(f1[p], f2[p], thr, 0.995, 1000, 12345)
                _same_as_oracle(o, g, p, N, (kinds[p], thr))
    finally:
        del os.environ["KML_EMU_STEWENIUS"]


def test_ransac_outcome_does_not_depend_on_the_round_schedule(oracle, ransacemu):
    """Latency schedule (128, 128, 256, 512 draws per round, small problem sets) against the throughput
    schedule (32, 32, 64, ...): the replay consumes draws in order, so iterations, winner, inlier set
    and model are the same, and the oracle's."""
    from test_gpu_parity import _nister_case
    rng = np.random.default_rng(8)
    f1, f2 = np.zeros((4, 50, 3)), np.zeros((4, 50, 3))
    for i, kind in enumerate(["plain", "low_parallax", "duplicates", "far_points"]):
        f1[i], f2[i] = _nister_case(rng, 50, kind)
    g0 = emu_ransac(ransacemu, True, f1, f2, 1e-6)
    os.environ["KML_EMU_LATENCY_SCHEDULE"] = "1"
    try:
        g1 = emu_ransac(ransacemu, True, f1, f2, 1e-6)
    finally:
        del os.environ["KML_EMU_LATENCY_SCHEDULE"]
    for k in ("iterations", "best_draw", "n_inliers", "mask", "models"):
        assert np.array_equal(g0[k], g1[k]), k
    for p in range(4):
        _same_as_oracle(oracle.ransac_nister(f1[p], f2[p], 1e-6, 0.995, 1000, 12345), g1, p, 50, p)
