"""kml — Python mirror of kimera_multi_lcd::LoopClosureDetector over libkml.so.

Method names, argument meaning and failure behaviour follow the reference class
(/root/reference/images/kimera-multi.drawio:2533-2609):
`addBowVector`, `addVLCFrame`, `detectLoopWithRobot`, `detectLoop`,
`computeMatchedIndices`, `geometricVerificationNister`, `recoverPose`,
`frameExists`, `bowExists`, `numBoWForRobot`, `getBoWVector`, counters
`totalBoWMatches` / `getNumGeomVerificationsMono` / `getNumGeomVerifications`.
All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/kml.h; this module only marshals numpy arrays through ctypes.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (KML_OK, KmlError, Params, Result, Stats, EXPORTS, build, lib,  # noqa: F401
                   Island, TemporalState, LCD_STATUS)

RESULT_DTYPE = np.dtype([
    ("q_robot", "<u8"), ("q_pose", "<u8"), ("m_robot", "<u8"), ("m_pose", "<u8"),
    ("norm_bow_score", "<f8"), ("n_matches", "<i4"), ("mono_inliers", "<i4"),
    ("stereo_inliers", "<i4"), ("status", "<i4"), ("R_mono", "<f8", (9,)), ("T", "<f8", (12,)),
])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


def default_params():
    p = Params()
    lib().kml_default_params(C.byref(p))
    return p


def params_from_yaml(path, literal_matcher_enum=False, base=None):
    """LcdParams.yaml (the reference's OpenCV-FileStorage parameter file,
    /root/reference/params/D455/LcdParams.yaml) -> Params, on top of `base` (default_params() if None).
    literal_matcher_enum: read matcher_type as cv::DescriptorMatcher::create() does (3 = L1) instead of
    as the file's own comment table documents it (3 = Hamming)."""
    p = base if base is not None else default_params()
    n = C.c_int(0)
    rc = lib().kml_params_from_yaml(str(path).encode(), int(bool(literal_matcher_enum)), C.byref(p), C.byref(n))
    if rc != KML_OK:
        raise KmlError(rc, (lib().kml_last_error(None) or b"").decode())
    p.n_mapped = n.value
    return p


def device_count():
    return lib().kml_device_count()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class _PinnedBlock:
    """page-locked host memory from kml_host_alloc, freed with the last array that views it"""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        rc = lib().kml_host_alloc(C.c_size_t(max(int(nbytes), 1)), C.byref(self.ptr))
        if rc != KML_OK:
            raise KmlError(rc, "kml_host_alloc(%d) failed" % nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib().kml_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (kml_host_alloc): batch arrays held like this are
    copied to the device from where they lie, without the library's staging copy."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    blk = _PinnedBlock(n * dtype.itemsize)
    buf = (C.c_uint8 * max(n * dtype.itemsize, 1)).from_address(blk.ptr.value)
    buf._kml_block = blk                       # keeps the block alive as long as any view exists
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def pinned_copy(a):
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class LoopClosureDetector:
    """B200-native stand-in for kimera_multi_lcd::LoopClosureDetector."""

    def __init__(self, params=None, device=0):
        self._h = C.c_void_p()
        self.params = params or default_params()
        self.loadAndInitialize(self.params, device)

    # ------------------------------------------------------------ lifecycle
    def loadAndInitialize(self, params, device=0):
        if self._h:
            lib().kml_destroy(self._h)
            self._h = C.c_void_p()
        rc = lib().kml_create(C.byref(params), int(device), C.byref(self._h))
        if rc != KML_OK:
            raise KmlError(rc, (lib().kml_last_error(None) or b"").decode())
        self.params = params
        self.device = device

    def create_lane(self):
        """Second query context on the same GPU sharing this detector's databases (kml_create_lane)."""
        lane = LoopClosureDetector.__new__(LoopClosureDetector)
        lane._h = C.c_void_p()
        lane.params, lane.device = self.params, self.device
        rc = lib().kml_create_lane(self._h, C.byref(lane._h))
        if rc != KML_OK:
            raise KmlError(rc, (lib().kml_last_error(self._h) or b"").decode())
        return lane

    def close(self):
        if getattr(self, "_h", None):
            lib().kml_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise KmlError(rc, (lib().kml_last_error(self._h) or b"").decode())
        return rc

    def stats(self):
        s = Stats()
        self._check(lib().kml_get_stats(self._h, C.byref(s)))
        return s

    def totalBoWMatches(self):
        return self.stats().total_bow_matches

    def getNumGeomVerificationsMono(self):
        return self.stats().total_geom_verifications_mono

    def getNumGeomVerifications(self):
        return self.stats().total_geometric_verifications

    # ------------------------------------------------------------ database
    def addBowVector(self, robot, pose, ids, vals):
        ids, vals = _u32(ids), _f32(vals)
        self._check(lib().kml_add_bow(self._h, C.c_uint64(robot), C.c_uint64(pose),
                                      _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids)))

    def addBowVectors(self, robot, poses, off, ids, vals):
        poses, off, ids, vals = _u64(poses), _i64(off), _u32(ids), _f32(vals)
        self._check(lib().kml_add_bow_bulk(self._h, C.c_uint64(robot), _p(poses, C.c_uint64),
                                           len(poses), _p(off, C.c_int64), _p(ids, C.c_uint32),
                                           _p(vals, C.c_float)))

    def addVLCFrame(self, robot, pose, desc, bearings, points):
        desc, bearings, points = _u8(desc), _f64(bearings), _f64(points)
        F = desc.size // 32
        self._check(lib().kml_add_frame(self._h, C.c_uint64(robot), C.c_uint64(pose),
                                        _p(desc, C.c_uint8), _p(bearings, C.c_double),
                                        _p(points, C.c_double), F))

    def addVLCFrames(self, robot, poses, desc, bearings, points):
        poses, desc, bearings, points = _u64(poses), _u8(desc), _f64(bearings), _f64(points)
        F = desc.size // 32 // max(len(poses), 1)
        self._check(lib().kml_add_frames_bulk(self._h, C.c_uint64(robot), _p(poses, C.c_uint64),
                                              len(poses), _p(desc, C.c_uint8),
                                              _p(bearings, C.c_double), _p(points, C.c_double), F))

    def frameExists(self, robot, pose):
        return bool(self._check(lib().kml_frame_exists(self._h, C.c_uint64(robot), C.c_uint64(pose))))

    def bowExists(self, robot, pose):
        return bool(self._check(lib().kml_bow_exists(self._h, C.c_uint64(robot), C.c_uint64(pose))))

    def numBoWForRobot(self, robot):
        return self._check(lib().kml_num_bow_for_robot(self._h, C.c_uint64(robot)))

    def getBoWVector(self, robot, pose, cap=4096):
        ids = np.zeros(cap, np.uint32)
        vals = np.zeros(cap, np.float32)
        cnt = C.c_int(0)
        rc = self._check(lib().kml_get_bow_vector(self._h, C.c_uint64(robot), C.c_uint64(pose),
                                                  _p(ids, C.c_uint32), _p(vals, C.c_float), cap,
                                                  C.byref(cnt)))
        if rc != KML_OK:
            return None
        return ids[:cnt.value].copy(), vals[:cnt.value].copy()

    def dbQuery(self, robot, ids, vals, max_results=1, max_id=-1):
        """DBoW2::TemplatedDatabase::query on one robot's database."""
        ids, vals = _u32(ids), _f32(vals)
        cap = max(int(max_results), 1)
        e = np.zeros(cap, np.uint32)
        s = np.zeros(cap, np.float64)
        cnt = C.c_int(0)
        self._check(lib().kml_db_query(self._h, C.c_uint64(robot), _p(ids, C.c_uint32),
                                       _p(vals, C.c_float), len(ids), int(max_results),
                                       int(max_id), _p(e, C.c_uint32), _p(s, C.c_double), cap,
                                       C.byref(cnt)))
        return e[:cnt.value].copy(), s[:cnt.value].copy()

    def score(self, ids1, vals1, ids2, vals2):
        """TemplatedVocabulary::score (L1)."""
        ids1, vals1, ids2, vals2 = _u32(ids1), _f32(vals1), _u32(ids2), _f32(vals2)
        out = C.c_double(0)
        self._check(lib().kml_bow_score(self._h, _p(ids1, C.c_uint32), _p(vals1, C.c_float),
                                        len(ids1), _p(ids2, C.c_uint32), _p(vals2, C.c_float),
                                        len(ids2), C.byref(out)))
        return out.value

    # ------------------------------------------------------------ detection
    def _detect(self, fn, head, ids, vals, cap):
        ids, vals = _u32(ids), _f32(vals)
        r = np.zeros(cap, np.uint64)
        p = np.zeros(cap, np.uint64)
        s = np.zeros(cap, np.float64)
        cnt = C.c_int(0)
        rc = self._check(fn(self._h, *head, _p(ids, C.c_uint32), _p(vals, C.c_float), len(ids),
                            _p(r, C.c_uint64), _p(p, C.c_uint64), _p(s, C.c_double), cap,
                            C.byref(cnt)))
        n = cnt.value
        return rc == KML_OK and n > 0, r[:n].copy(), p[:n].copy(), s[:n].copy()

    def detectLoopWithRobot(self, robot, q_robot, q_pose, ids, vals, cap=128):
        return self._detect(lib().kml_detect_loop_with_robot,
                            (C.c_uint64(robot), C.c_uint64(q_robot), C.c_uint64(q_pose)),
                            ids, vals, cap)

    def detectLoop(self, q_robot, q_pose, ids, vals, cap=2048):
        return self._detect(lib().kml_detect_loop, (C.c_uint64(q_robot), C.c_uint64(q_pose)),
                            ids, vals, cap)

    # --------------------------------------------------------- verification
    def computeMatchedIndices(self, qr, qp, mr, mp, cap=65536):
        iq = np.zeros(cap, np.uint32)
        im = np.zeros(cap, np.uint32)
        cnt = C.c_int(0)
        self._check(lib().kml_compute_matched_indices(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), cap, C.byref(cnt)))
        return iq[:cnt.value].copy(), im[:cnt.value].copy()

    def geometricVerificationNister(self, qr, qp, mr, mp, inl_q, inl_m):
        iq, im = _u32(inl_q).copy(), _u32(inl_m).copy()
        cnt = C.c_int(len(iq))
        if len(iq) == 0:
            iq, im = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        R = np.zeros((3, 3))
        rc = self._check(lib().kml_geometric_verification_nister(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), C.byref(cnt), _p(R, C.c_double)))
        return rc == KML_OK, iq[:cnt.value].copy(), im[:cnt.value].copy(), R

    def recoverPose(self, qr, qp, mr, mp, inl_q, inl_m, R_prior=None):
        iq, im = _u32(inl_q).copy(), _u32(inl_m).copy()
        cnt = C.c_int(len(iq))
        if len(iq) == 0:
            iq, im = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        T = np.zeros((3, 4))
        pr = None
        if R_prior is not None:
            R_prior = _f64(R_prior)
            pr = _p(R_prior, C.c_double)
        rc = self._check(lib().kml_recover_pose(
            self._h, C.c_uint64(qr), C.c_uint64(qp), C.c_uint64(mr), C.c_uint64(mp),
            _p(iq, C.c_uint32), _p(im, C.c_uint32), C.byref(cnt), pr, _p(T, C.c_double)))
        return rc == KML_OK, iq[:cnt.value].copy(), im[:cnt.value].copy(), T

    # -------------------------------------------------------- batched paths
    def _batch_args(self, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                    desc, bearings, points):
        B = len(q_robot)
        keep = (_u64(q_robot), _u64(q_pose), _i64(bow_off), _u32(ids), _f32(vals), _i64(prev_off),
                _u32(prev_ids), _f32(prev_vals), _u8(desc), _f64(bearings), _f64(points))
        F = keep[8].size // 32 // max(B, 1)
        args = (B, _p(keep[0], C.c_uint64), _p(keep[1], C.c_uint64), _p(keep[2], C.c_int64),
                _p(keep[3], C.c_uint32), _p(keep[4], C.c_float), _p(keep[5], C.c_int64),
                _p(keep[6], C.c_uint32), _p(keep[7], C.c_float), _p(keep[8], C.c_uint8),
                _p(keep[9], C.c_double), _p(keep[10], C.c_double), F)
        return keep, args

    def query_batch(self, q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids, prev_vals,
                    desc, bearings, points, sharded=False, seq=None):
        """B full loop-closure queries, host buffers in, records out (kml_query_batch; with
        sharded=True kml_query_batch_sharded_host: every rank passes the same batch)."""
        keep, args = self._batch_args(q_robot, q_pose, bow_off, ids, vals, prev_off, prev_ids,
                                      prev_vals, desc, bearings, points)
        B = args[0]
        cap = int(self.params.top_k_verify)
        out = np.zeros((B, cap), RESULT_DTYPE)
        counts = np.zeros(B, np.int32)
        if sharded:
            self._check(lib().kml_query_batch_sharded_host(self._h, int(seq is not None), C.c_uint64(int(seq or 0)), *args,
                                                           out.ctypes.data_as(C.c_void_p), cap, _p(counts, C.c_int32)))
            return out, counts
        self._check(lib().kml_query_batch(self._h, *args, out.ctypes.data_as(C.c_void_p), cap,
                                          _p(counts, C.c_int32)))
        return out, counts

    def query_batch_upload(self, *a):
        keep, args = self._batch_args(*a)
        self._check(lib().kml_query_batch_upload(self._h, *args))
        self._B = args[0]

    def query_batch_run(self, sharded=False, seq=None):
        """Run the uploaded batch.  sharded: merge with the other ranks' shards (one ncclAllGather);
        seq: global sequence number of this batch when several lanes of a rank exchange
        concurrently (their collectives are then enqueued in seq order on every rank)."""
        B = self._B
        cap = int(self.params.top_k_verify)
        out = np.zeros((B, cap), RESULT_DTYPE)
        counts = np.zeros(B, np.int32)
        if sharded and seq is not None:
            self._check(lib().kml_query_batch_sharded_seq(self._h, C.c_uint64(int(seq)), out.ctypes.data_as(C.c_void_p),
                                                          cap, _p(counts, C.c_int32)))
            return out, counts
        fn = lib().kml_query_batch_sharded if sharded else lib().kml_query_batch_run
        self._check(fn(self._h, out.ctypes.data_as(C.c_void_p), cap, _p(counts, C.c_int32)))
        return out, counts

    def comm_seq_reset(self, next_seq=0):
        """next sequence number kml_query_batch_sharded_seq waits for (shared by this detector's lanes)"""
        self._check(lib().kml_comm_seq_reset(self._h, C.c_uint64(int(next_seq))))

    def merge_shard_records_device(self, blocks, cap):
        """merge_shards_kernel on host blocks [(records[B, cap_in], counts[B]), ...] (test hook of the
        sharded query's device tail)"""
        B = len(blocks[0][1])
        cap_in = blocks[0][0].shape[1]
        buf = b"".join(np.ascontiguousarray(rec).tobytes() + np.ascontiguousarray(cnt, np.int32).tobytes()
                       for rec, cnt in blocks)
        out = np.zeros((B, cap), RESULT_DTYPE)
        counts = np.zeros(B, np.int32)
        self._check(lib().kml_merge_shard_records_device(self._h, buf, len(blocks), B, cap_in, cap,
                                                         out.ctypes.data_as(C.c_void_p), _p(counts, C.c_int32)))
        return out, counts

    def hamming_knn2(self, q, t, reps=0):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) -> idx[nq,2], dist[nq,2], ms."""
        q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.uint32)
        dist = np.zeros((len(q), 2), np.uint16)
        ms = C.c_float(0)
        if reps > 0:
            self._check(lib().kml_hamming_knn2_bench(
                self._h, _p(q, C.c_uint8), len(q), _p(t, C.c_uint8), C.c_int64(len(t)), reps,
                _p(idx, C.c_uint32), _p(dist, C.c_uint16), C.byref(ms)))
        else:
            self._check(lib().kml_hamming_knn2(
                self._h, _p(q, C.c_uint8), len(q), _p(t, C.c_uint8), C.c_int64(len(t)),
                _p(idx, C.c_uint32), _p(dist, C.c_uint16), C.byref(ms)))
        return idx, dist, ms.value

    def l1_knn2(self, q, t):
        """cv::BFMatcher(NORM_L1).knnMatch(q, t, k=2) on descriptor bytes -> idx[nq,2], dist[nq,2], ms."""
        q, t = _u8(q).reshape(-1, 32), _u8(t).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.uint32)
        dist = np.zeros((len(q), 2), np.uint16)
        ms = C.c_float(0)
        self._check(lib().kml_l1_knn2(self._h, _p(q, C.c_uint8), len(q), _p(t, C.c_uint8), C.c_int64(len(t)),
                                      _p(idx, C.c_uint32), _p(dist, C.c_uint16), C.byref(ms)))
        return idx, dist, ms.value

    def _ransac_batch(self, fn, a, b, full):
        a, b = _f64(a), _f64(b)
        P, N = a.shape[0], a.shape[1]
        models = np.zeros((P, 3, 4))
        n_inl = np.zeros(P, np.int32)
        iters = np.zeros(P, np.int32)
        best = np.zeros(P, np.int32)
        words = (N + 31) // 32
        mask = np.zeros((P, max(words, 1)), np.uint32)
        ms = C.c_float(0)
        self._check(fn(self._h, P, N, _p(a, C.c_double), _p(b, C.c_double), int(full),
                       _p(models, C.c_double), _p(n_inl, C.c_int32), _p(iters, C.c_int32),
                       _p(best, C.c_int32), _p(mask, C.c_uint32), C.byref(ms)))
        return dict(models=models, n_inliers=n_inl, iterations=iters, best_draw=best, mask=mask,
                    ms=ms.value)

    def ransac_arun_batch(self, p1, p2, full_hypotheses=False):
        return self._ransac_batch(lib().kml_ransac_arun_batch, p1, p2, full_hypotheses)

    def ransac_onepoint_batch(self, p1, p2, R, full_hypotheses=False):
        """1-point stereo RANSAC with the rotations R [P, 3, 3] given (row f4)."""
        R = _f64(R).reshape(-1, 9)
        fn = lib().kml_ransac_onepoint_batch

        def call(h, P, N, a, b, full, *rest):
            return fn(h, P, N, a, b, _p(R, C.c_double), full, *rest)
        return self._ransac_batch(call, p1, p2, full_hypotheses)

    def ransac_nister_batch(self, f1, f2, full_hypotheses=False):
        return self._ransac_batch(lib().kml_ransac_nister_batch, f1, f2, full_hypotheses)

    def peak_popc(self):
        out = C.c_double(0)
        self._check(lib().kml_peak_popc(self._h, C.byref(out)))
        return out.value

    # ------------------------------------------------ vocabulary (row f1)
    def setVocabulary(self, k, L, node_desc, word_weights):
        """OrbVocabulary: k-ary tree of L levels, nodes breadth-first, IDF weight per leaf (= word)."""
        node_desc, word_weights = _u8(node_desc), _f64(word_weights)
        self._check(lib().kml_vocab_set(self._h, int(k), int(L), _p(node_desc, C.c_uint8),
                                        _p(word_weights, C.c_double)))

    def transform(self, desc):
        """TemplatedVocabulary::transform for [B, F, 32] descriptors -> (off[B+1], ids, vals, ms)."""
        desc = _u8(desc)
        B, F = desc.shape[0], desc.shape[1]
        off = np.zeros(B + 1, np.int64)
        ids = np.zeros(max(B * F, 1), np.uint32)
        vals = np.zeros(max(B * F, 1), np.float64)
        ms = C.c_float(0)
        self._check(lib().kml_transform_batch(self._h, B, F, _p(desc, C.c_uint8), _p(off, C.c_int64),
                                              _p(ids, C.c_uint32), _p(vals, C.c_double), C.c_int64(B * F),
                                              C.byref(ms)))
        return off, ids[:off[-1]].copy(), vals[:off[-1]].copy(), ms.value

    def timer_begin(self):
        self._check(lib().kml_timer_begin(self._h))

    def timer_end(self, lanes=()):
        arr = (C.c_void_p * max(len(lanes), 1))(*[l._h for l in lanes])
        ms = C.c_float(0)
        self._check(lib().kml_timer_end(self._h, arr, len(lanes), C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self._check(lib().kml_flush_l2(self._h))

    def peak_fp64(self):
        out = C.c_double(0)
        self._check(lib().kml_peak_fp64(self._h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------ multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * _lib.KML_UNIQUE_ID_BYTES)()
        rc = lib().kml_comm_unique_id(buf)
        if rc != KML_OK:
            raise KmlError(rc, "kml_comm_unique_id failed")
        return bytes(buf)

    # ------------------------------------------------ row f4: shard persistence
    def save(self, path):
        """Write this detector's BoW databases and frames to one file (kml_save_shard)."""
        self._check(lib().kml_save_shard(self._h, str(path).encode()))

    def load(self, path):
        """Replay a saved shard into this (empty) detector (kml_load_shard)."""
        self._check(lib().kml_load_shard(self._h, str(path).encode()))

    # ------------------------------------------------ row f3: post filters, wire layout
    def addVLCFrameMsg(self, robot, pose, desc, versors_f32, keypoints_f32):
        """addVLCFrame from the VLCFrameMsg layout (float32 xyz clouds)."""
        desc = _u8(desc)
        v = np.ascontiguousarray(versors_f32, np.float32)
        k = np.ascontiguousarray(keypoints_f32, np.float32)
        self._check(lib().kml_add_frame_msg(self._h, C.c_uint64(int(robot)), C.c_uint64(int(pose)),
                                            _p(desc, C.c_uint8), _p(v, C.c_float), _p(k, C.c_float),
                                            int(desc.shape[0])))

    def detectLoopIslands(self, robot, q_robot, q_pose, ids, vals, state, max_intraisland_gap=3,
                          min_matches_per_island=1, max_nrFrames_between_islands=3, min_temporal_matches=1):
        """detectLoopWithRobot + computeIslands + checkTemporalConstraint (Kimera-VIO flow).
        Returns (lcd_status string, match_pose, match_score, island tuple or None)."""
        ids = np.ascontiguousarray(ids, np.uint32)
        vals = np.ascontiguousarray(vals, np.float32)
        mp, ms, isl, st = C.c_uint64(0), C.c_double(0), Island(), C.c_int(1)
        rc = lib().kml_detect_loop_islands(self._h, C.c_uint64(int(robot)), C.c_uint64(int(q_robot)),
                                           C.c_uint64(int(q_pose)), _p(ids, C.c_uint32), _p(vals, C.c_float),
                                           len(ids), int(max_intraisland_gap), int(min_matches_per_island),
                                           int(max_nrFrames_between_islands), int(min_temporal_matches),
                                           C.byref(state), C.byref(mp), C.byref(ms), C.byref(isl), C.byref(st))
        if rc < 0:
            self._check(rc)
        status = LCD_STATUS[st.value]
        has_island = status in ("LOOP_DETECTED", "FAILED_TEMPORAL_CONSTRAINT")
        return status, int(mp.value), float(ms.value), (isl.astuple() if has_island else None)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_uint8 * _lib.KML_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(lib().kml_comm_init(self._h, int(nranks), int(rank), buf))


def compute_islands(ids, scores, max_intraisland_gap=3, min_matches_per_island=1):
    """LcdThirdPartyWrapper::computeIslands on one database's surviving results -> list of
    (start_id, end_id, best_id, island_score, best_score), ascending start id."""
    ids = np.ascontiguousarray(ids, np.uint64)
    scores = np.ascontiguousarray(scores, np.float64)
    out = (Island * max(len(ids), 1))()
    cnt = C.c_int(0)
    rc = lib().kml_compute_islands(_p(ids, C.c_uint64), _p(scores, C.c_double), len(ids), int(max_intraisland_gap),
                                   int(min_matches_per_island), out, len(out), C.byref(cnt))
    if rc != KML_OK:
        raise KmlError(rc, "kml_compute_islands")
    return [out[i].astuple() for i in range(cnt.value)]


def check_temporal_constraint(state, query_id, island, max_nrFrames_between_queries=2,
                              max_nrFrames_between_islands=3, min_temporal_matches=1):
    """LcdThirdPartyWrapper::checkTemporalConstraint; `state` is a kml.TemporalState()."""
    isl = Island(*island)
    rc = lib().kml_check_temporal_constraint(C.byref(state), C.c_uint64(int(query_id)), C.byref(isl),
                                             int(max_nrFrames_between_queries), int(max_nrFrames_between_islands),
                                             int(min_temporal_matches))
    if rc < 0:
        raise KmlError(rc, "kml_check_temporal_constraint")
    return bool(rc)


def mask_to_indices(mask_row, n):
    """inlier bitmask words -> ascending index array"""
    bits = np.unpackbits(mask_row.view(np.uint8), bitorder="little")[:n]
    return np.nonzero(bits)[0].astype(np.uint32)


from . import logio  # noqa: E402,F401
