"""ctypes binding of the CPU ORACLE (oracle/libkml_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(kimera-multi_b200/kml) never imports this module.  See oracle/kmo.h for the
parity status ("parity unpinned" against the un-vendored upstream libraries;
pinned against cv2 / numpy / SURVEY.md Appendix B known answers instead).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkml_oracle.so")


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [
        ("inter_robot_only", C.c_int32),
        ("alpha", C.c_double),
        ("dist_local", C.c_int32),
        ("max_db_results", C.c_int32),
        ("min_nss_factor", C.c_double),
        ("max_nrFrames_between_queries", C.c_int32),
        ("lowe_ratio", C.c_double),
        ("ransac_threshold_mono", C.c_double),
        ("ransac_inlier_percentage_mono", C.c_double),
        ("max_ransac_iterations_mono", C.c_int32),
        ("ransac_probability_mono", C.c_double),
        ("ransac_threshold", C.c_double),
        ("max_ransac_iterations", C.c_int32),
        ("ransac_probability", C.c_double),
        ("geometric_verification_min_inlier_count", C.c_int32),
        ("geometric_verification_min_inlier_percentage", C.c_double),
        ("ransac_seed", C.c_uint32),
        ("top_k_verify", C.c_int32),
        ("matcher_norm", C.c_int32),
        ("mono_algorithm", C.c_int32),
        ("ransac_use_1point_3d3d", C.c_int32),
    ]


class RansacResult(C.Structure):
    _fields_ = [
        ("success", C.c_int32),
        ("iterations", C.c_int32),
        ("skipped", C.c_int32),
        ("draws", C.c_int32),
        ("best_draw", C.c_int32),
        ("n_inliers", C.c_int32),
        ("model", C.c_double * 12),
    ]


class Result(C.Structure):
    _fields_ = [
        ("q_robot", C.c_uint64), ("q_pose", C.c_uint64),
        ("m_robot", C.c_uint64), ("m_pose", C.c_uint64),
        ("norm_bow_score", C.c_double),
        ("n_matches", C.c_int32), ("mono_inliers", C.c_int32),
        ("stereo_inliers", C.c_int32), ("status", C.c_int32),
        ("R_mono", C.c_double * 9),
        ("T", C.c_double * 12),
    ]


RESULT_DTYPE = np.dtype([
    ("q_robot", "<u8"), ("q_pose", "<u8"), ("m_robot", "<u8"), ("m_pose", "<u8"),
    ("norm_bow_score", "<f8"), ("n_matches", "<i4"), ("mono_inliers", "<i4"),
    ("stereo_inliers", "<i4"), ("status", "<i4"), ("R_mono", "<f8", (9,)), ("T", "<f8", (12,)),
])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)

_lib = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.kmo_bow_score.restype = C.c_double
        L.kmo_db_create.restype = C.c_void_p
        L.kmo_db_add.restype = C.c_uint32
        L.kmo_db_size.restype = C.c_uint32
        L.kmo_lcd_create.restype = C.c_void_p
        L.kmo_mono_residual.restype = C.c_double
        L.kmo_arun_residual.restype = C.c_double
        _lib = L
    return _lib


def default_params():
    p = Params()
    lib().kmo_default_params(C.byref(p))
    return p


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def bow_score(ids1, vals1, ids2, vals2):
    ids1, vals1, ids2, vals2 = _u32(ids1), _f32(vals1), _u32(ids2), _f32(vals2)
    return lib().kmo_bow_score(_p(ids1, C.c_uint32), _p(vals1, C.c_float), len(ids1),
                               _p(ids2, C.c_uint32), _p(vals2, C.c_float), len(ids2))


class Database:
    """DBoW2::TemplatedDatabase restatement."""

    def __init__(self):
        self._h = C.c_void_p(lib().kmo_db_create())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmo_db_destroy(self._h)
            self._h = None

    def add(self, ids, vals):
        ids, vals = _u32(ids), _f32(vals)
        return lib().kmo_db_add(self._h, _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids))

    def size(self):
        return lib().kmo_db_size(self._h)

    def query(self, ids, vals, max_results=1, max_id=-1):
        ids, vals = _u32(ids), _f32(vals)
        cap = max_results if max_results > 0 else max(self.size(), 1)
        e = np.zeros(cap, np.uint32)
        s = np.zeros(cap, np.float64)
        n = lib().kmo_db_query(self._h, _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids),
                               max_results, max_id, _p(e, C.c_uint32), _p(s, C.c_double), cap)
        return e[:n].copy(), s[:n].copy()


def hamming_knn2(q, t):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    idx = np.zeros((len(q), 2), np.uint32)
    dist = np.zeros((len(q), 2), np.uint16)
    lib().kmo_hamming_knn2(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t),
                           _p(idx, C.c_uint32), _p(dist, C.c_uint16))
    return idx, dist


def l1_knn2(q, t):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    idx = np.zeros((len(q), 2), np.uint32)
    dist = np.zeros((len(q), 2), np.uint16)
    lib().kmo_l1_knn2(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(idx, C.c_uint32), _p(dist, C.c_uint16))
    return idx, dist


def match_lowe(q, t, lowe_ratio):
    q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
    iq = np.zeros(max(len(q), 1), np.uint32)
    im = np.zeros(max(len(q), 1), np.uint32)
    n = lib().kmo_match_lowe(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t),
                             C.c_double(lowe_ratio), _p(iq, C.c_uint32), _p(im, C.c_uint32))
    return iq[:n].copy(), im[:n].copy()


def sample_stream(N, s, seed, n_draws):
    out = np.zeros((n_draws, s), np.uint16)
    lib().kmo_sample_stream(N, s, C.c_uint32(seed), n_draws, _p(out, C.c_uint16))
    return out


def arun3(p1, p2):
    p1, p2 = _f64(p1).reshape(3, 3), _f64(p2).reshape(3, 3)
    m = np.zeros(12)
    lib().kmo_arun3(_p(p1, C.c_double), _p(p2, C.c_double), _p(m, C.c_double))
    return m.reshape(3, 4)


def svd3(A):
    A = _f64(A).reshape(3, 3)
    U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    lib().kmo_svd3(_p(A, C.c_double), _p(U, C.c_double), _p(S, C.c_double), _p(V, C.c_double))
    return U, S, V


def fivept_nister(f1, f2):
    f1, f2 = _f64(f1).reshape(5, 3), _f64(f2).reshape(5, 3)
    E = np.zeros((10, 3, 3))
    n = lib().kmo_fivept_nister(_p(f1, C.c_double), _p(f2, C.c_double), _p(E, C.c_double))
    return E[:n].copy()


def fivept_stewenius(f1, f2):
    """row f4: the Stewenius solver's real solutions (same conventions as fivept_nister)"""
    f1, f2 = _f64(f1).reshape(5, 3), _f64(f2).reshape(5, 3)
    E = np.zeros((10, 3, 3))
    n = lib().kmo_fivept_stewenius(_p(f1, C.c_double), _p(f2, C.c_double), _p(E, C.c_double))
    return E[:n].copy()


def mono_model_alg(f1, f2, sample8, algorithm):
    f1, f2 = _f64(f1).reshape(-1, 3), _f64(f2).reshape(-1, 3)
    s8 = np.ascontiguousarray(sample8, np.uint16)
    m = np.zeros(12)
    ok = lib().kmo_mono_model_alg(_p(f1, C.c_double), _p(f2, C.c_double), _p(s8, C.c_uint16), _p(m, C.c_double), int(algorithm))
    return bool(ok), m.reshape(3, 4)


def mono_model(f1, f2, sample8):
    f1, f2 = _f64(f1).reshape(-1, 3), _f64(f2).reshape(-1, 3)
    s = np.ascontiguousarray(sample8, dtype=np.uint16)
    m = np.zeros(12)
    ok = lib().kmo_mono_model(_p(f1, C.c_double), _p(f2, C.c_double), _p(s, C.c_uint16),
                              _p(m, C.c_double))
    return bool(ok), m.reshape(3, 4)


def mono_residual(model, f1, f2):
    model, f1, f2 = _f64(model).reshape(12), _f64(f1), _f64(f2)
    return lib().kmo_mono_residual(_p(model, C.c_double), _p(f1, C.c_double), _p(f2, C.c_double))


def arun_residual(model, p1, p2):
    model, p1, p2 = _f64(model).reshape(12), _f64(p1), _f64(p2)
    return lib().kmo_arun_residual(_p(model, C.c_double), _p(p1, C.c_double), _p(p2, C.c_double))


def _ransac(fn, a, b, thr, prob, max_iter, seed):
    a, b = _f64(a).reshape(-1, 3), _f64(b).reshape(-1, 3)
    res = RansacResult()
    inl = np.zeros(max(len(a), 1), np.uint32)
    fn(_p(a, C.c_double), _p(b, C.c_double), len(a), C.c_double(thr), C.c_double(prob),
       int(max_iter), C.c_uint32(seed), C.byref(res), _p(inl, C.c_uint32))
    return dict(success=bool(res.success), iterations=res.iterations, skipped=res.skipped,
                draws=res.draws, best_draw=res.best_draw, n_inliers=res.n_inliers,
                model=np.array(res.model).reshape(3, 4),
                inliers=inl[:res.n_inliers].copy() if res.success else np.zeros(0, np.uint32))


def ransac_arun(p1, p2, thr=0.5, prob=0.995, max_iter=1000, seed=12345):
    return _ransac(lib().kmo_ransac_arun, p1, p2, thr, prob, max_iter, seed)


def ransac_onepoint(p1, p2, R, thr=0.5, prob=0.995, max_iter=1000, seed=12345):
    """Ransac over the 1-point stereo problem: rotation R given, model = [R | p1_i - R p2_i] (row f4)."""
    a, b, R = _f64(p1).reshape(-1, 3), _f64(p2).reshape(-1, 3), _f64(R).reshape(9)
    res = RansacResult()
    inl = np.zeros(max(len(a), 1), np.uint32)
    lib().kmo_ransac_onepoint(_p(a, C.c_double), _p(b, C.c_double), len(a), _p(R, C.c_double), C.c_double(thr),
                              C.c_double(prob), int(max_iter), C.c_uint32(seed), C.byref(res), _p(inl, C.c_uint32))
    return dict(success=bool(res.success), iterations=res.iterations, skipped=res.skipped,
                draws=res.draws, best_draw=res.best_draw, n_inliers=res.n_inliers,
                model=np.array(res.model).reshape(3, 4),
                inliers=inl[:res.n_inliers].copy() if res.success else np.zeros(0, np.uint32))


def ransac_nister(f1, f2, thr=1e-6, prob=0.995, max_iter=1000, seed=12345):
    return _ransac(lib().kmo_ransac_nister, f1, f2, thr, prob, max_iter, seed)


def debug_root_grid2(on):
    """test knob: 0 = skip the 256-cell root grid (pairs with the library's KML_NO_ROOT_GRID2)"""
    lib().kmo_debug_root_grid2(C.c_int(int(on)))


def krsqrt(x):
    """the contract's reciprocal square root, elementwise"""
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    lib().kmo_krsqrt(_p(x, C.c_double), C.c_int(x.size), _p(out, C.c_double))
    return out


def ransac_stewenius(f1, f2, thr=1e-6, prob=0.995, max_iter=1000, seed=12345):
    return _ransac(lib().kmo_ransac_stewenius, f1, f2, thr, prob, max_iter, seed)


class LoopClosureDetector:
    """kimera_multi_lcd::LoopClosureDetector restatement."""

    def __init__(self, params=None):
        self.params = params or default_params()
        self._h = C.c_void_p(lib().kmo_lcd_create(C.byref(self.params)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmo_lcd_destroy(self._h)
            self._h = None

    def addBowVector(self, robot, pose, ids, vals):
        ids, vals = _u32(ids), _f32(vals)
        lib().kmo_lcd_add_bow(self._h, C.c_uint64(robot), C.c_uint64(pose),
                              _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids))

    def addVLCFrame(self, robot, pose, desc, bearings, points):
        desc, bearings, points = _u8(desc), _f64(bearings), _f64(points)
        F = desc.size // 32
        lib().kmo_lcd_add_frame(self._h, C.c_uint64(robot), C.c_uint64(pose), _p(desc, C.c_uint8),
                                _p(bearings, C.c_double), _p(points, C.c_double), F)

    def _detect(self, fn, head, ids, vals, cap):
        ids, vals = _u32(ids), _f32(vals)
        r = np.zeros(cap, np.uint64)
        p = np.zeros(cap, np.uint64)
        s = np.zeros(cap, np.float64)
        n = fn(self._h, *head, _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids),
               _p(r, C.c_uint64), _p(p, C.c_uint64), _p(s, C.c_double), cap)
        return r[:n].copy(), p[:n].copy(), s[:n].copy()

    def detectLoopWithRobot(self, robot, q_robot, q_pose, ids, vals, cap=64):
        return self._detect(lib().kmo_lcd_detect_loop_with_robot,
                            (C.c_uint64(robot), C.c_uint64(q_robot), C.c_uint64(q_pose)),
                            ids, vals, cap)

    def detectLoop(self, q_robot, q_pose, ids, vals, cap=1024):
        return self._detect(lib().kmo_lcd_detect_loop,
                            (C.c_uint64(q_robot), C.c_uint64(q_pose)), ids, vals, cap)

    def computeMatchedIndices(self, qr, qp, mr, mp, cap=4096):
        iq = np.zeros(cap, np.uint32)
        im = np.zeros(cap, np.uint32)
        n = lib().kmo_lcd_compute_matched_indices(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), cap)
        return iq[:n].copy(), im[:n].copy()

    def geometricVerificationNister(self, qr, qp, mr, mp, inl_q, inl_m):
        iq, im = _u32(inl_q).copy(), _u32(inl_m).copy()
        cnt = C.c_int(len(iq))
        R = np.zeros((3, 3))
        if len(iq) == 0:
            iq, im = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        ok = lib().kmo_lcd_geometric_verification_nister(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), C.byref(cnt), _p(R, C.c_double))
        return bool(ok), iq[:cnt.value].copy(), im[:cnt.value].copy(), R

    def recoverPose(self, qr, qp, mr, mp, inl_q, inl_m, R_prior=None):
        iq, im = _u32(inl_q).copy(), _u32(inl_m).copy()
        cnt = C.c_int(len(iq))
        T = np.zeros((3, 4))
        if len(iq) == 0:
            iq, im = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        pr = None
        if R_prior is not None:
            R_prior = _f64(R_prior)
            pr = _p(R_prior, C.c_double)
        ok = lib().kmo_lcd_recover_pose_prior(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), C.byref(cnt), pr, _p(T, C.c_double))
        return bool(ok), iq[:cnt.value].copy(), im[:cnt.value].copy(), T

    def query_batch(self, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                    desc, bearings, points, threads=0):
        """Full loop-closure queries (see kmo_lcd_query in oracle/kmo.h)."""
        B = len(q_robot)
        q_robot = np.ascontiguousarray(q_robot, np.uint64)
        q_pose = np.ascontiguousarray(q_pose, np.uint64)
        bow_off = np.ascontiguousarray(bow_off, np.int64)
        prev_off = np.ascontiguousarray(prev_off, np.int64)
        ids, vals, prev_ids, prev_vals = _u32(ids), _f32(vals), _u32(prev_ids), _f32(prev_vals)
        desc, bearings, points = _u8(desc), _f64(bearings), _f64(points)
        F = desc.size // 32 // max(B, 1)
        cap = int(self.params.top_k_verify)
        out = np.zeros((B, cap), RESULT_DTYPE)
        counts = np.zeros(B, np.int32)
        lib().kmo_lcd_query_batch(
            self._h, B, _p(q_robot, C.c_uint64), _p(q_pose, C.c_uint64), _p(bow_off, C.c_int64),
            _p(ids, C.c_uint32), _p(vals, C.c_float), _p(prev_off, C.c_int64),
            _p(prev_ids, C.c_uint32), _p(prev_vals, C.c_float), _p(desc, C.c_uint8),
            _p(bearings, C.c_double), _p(points, C.c_double), F,
            out.ctypes.data_as(C.c_void_p), cap, _p(counts, C.c_int32), int(threads))
        return out, counts


class Vocabulary:
    """DBoW2::TemplatedVocabulary::transform restatement."""

    def __init__(self, k, L, node_desc, word_weights):
        node_desc, word_weights = _u8(node_desc), _f64(word_weights)
        lib().kmo_vocab_create.restype = C.c_void_p
        self._h = C.c_void_p(lib().kmo_vocab_create(int(k), int(L), _p(node_desc, C.c_uint8),
                                                    _p(word_weights, C.c_double)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kmo_vocab_destroy(self._h)
            self._h = None

    def transform(self, desc):
        desc = _u8(desc).reshape(-1, 32)
        ids = np.zeros(max(len(desc), 1), np.uint32)
        vals = np.zeros(max(len(desc), 1), np.float64)
        n = lib().kmo_vocab_transform(self._h, _p(desc, C.c_uint8), len(desc), _p(ids, C.c_uint32),
                                      _p(vals, C.c_double))
        return ids[:n].copy(), vals[:n].copy()


def num_threads():
    return lib().kmo_num_threads()


# ---------------------------------------------------------------- row f3
class Island(C.Structure):
    _fields_ = [("start_id", C.c_uint64), ("end_id", C.c_uint64), ("best_id", C.c_uint64),
                ("island_score", C.c_double), ("best_score", C.c_double)]

    def astuple(self):
        return (self.start_id, self.end_id, self.best_id, self.island_score, self.best_score)


class TemporalState(C.Structure):
    _fields_ = [("temporal_entries", C.c_int32), ("pad", C.c_int32), ("latest_query_id", C.c_uint64),
                ("latest_island", Island)]


def compute_islands(ids, scores, max_gap, min_len):
    ids = np.ascontiguousarray(ids, np.uint64)
    scores = _f64(scores)
    out = (Island * max(len(ids), 1))()
    lib().kmo_compute_islands.restype = C.c_int
    n = lib().kmo_compute_islands(_p(ids, C.c_uint64), _p(scores, C.c_double), len(ids), int(max_gap),
                                  int(min_len), out, len(out))
    return [out[i].astuple() for i in range(n)]


def check_temporal_constraint(state, query_id, island, max_between_queries, max_between_islands,
                              min_temporal_matches):
    isl = Island(*island)
    lib().kmo_check_temporal_constraint.restype = C.c_int
    return bool(lib().kmo_check_temporal_constraint(C.byref(state), C.c_uint64(int(query_id)), C.byref(isl),
                                                    int(max_between_queries), int(max_between_islands),
                                                    int(min_temporal_matches)))
