"""ctypes loader for libkml.so (the sm_100a product library).

No PyTorch, no fallback: if the shared library is missing or no B200-class GPU
is usable, loading / kml_create fails loudly.  Mirrors include/kml.h.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libkml.so")

KML_OK = 0
KML_NO_DB, KML_NO_PREV_BOW, KML_NSS_TOO_LOW, KML_NO_MATCH, KML_NO_FRAME = 1, 2, 3, 4, 5
KML_TOO_FEW_POINTS, KML_RANSAC_FAIL, KML_TOO_FEW_INLIERS, KML_INTER_ROBOT_ONLY = 6, 7, 8, 9
KML_ERR_ARG, KML_ERR_CUDA, KML_ERR_NCCL, KML_ERR_CAPACITY, KML_ERR_STREAM_EXHAUSTED, KML_ERR_IO = -1, -2, -3, -4, -5, -6
KML_UNIQUE_ID_BYTES = 128


class KmlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libkml error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("inter_robot_only", C.c_int32),
        ("alpha", C.c_double),
        ("dist_local", C.c_int32),
        ("max_db_results", C.c_int32),
        ("min_nss_factor", C.c_double),
        ("max_nrFrames_between_queries", C.c_int32),
        ("max_nrFrames_between_islands", C.c_int32),
        ("min_temporal_matches", C.c_int32),
        ("max_intraisland_gap", C.c_int32),
        ("min_matches_per_island", C.c_int32),
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
        ("ransac_randomize", C.c_int32),
        ("ransac_seed", C.c_uint32),
        ("top_k_verify", C.c_int32),
        ("matcher_norm", C.c_int32),
        ("matcher_engine", C.c_int32),
        ("mono_algorithm", C.c_int32),
        ("ransac_use_1point_3d3d", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("total_bow_matches", C.c_uint64),
        ("total_geom_verifications_mono", C.c_uint64),
        ("total_geometric_verifications", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("ms_bow", C.c_float), ("ms_match", C.c_float), ("ms_mono", C.c_float),
        ("ms_stereo", C.c_float), ("ms_total", C.c_float),
        ("bow_postings_last", C.c_uint64),
        ("mono_hypotheses_last", C.c_uint64), ("stereo_hypotheses_last", C.c_uint64),
        ("pairs_last", C.c_uint64),
        ("mono_residuals_last", C.c_uint64), ("stereo_residuals_last", C.c_uint64),
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


# every symbol include/kml.h declares (checked by tests/test_abi.py)
class Island(C.Structure):
    """kml_island"""
    _fields_ = [("start_id", C.c_uint64), ("end_id", C.c_uint64), ("best_id", C.c_uint64),
                ("island_score", C.c_double), ("best_score", C.c_double)]

    def astuple(self):
        return (self.start_id, self.end_id, self.best_id, self.island_score, self.best_score)


class TemporalState(C.Structure):
    """kml_temporal_state (zero-initialised = no history)"""
    _fields_ = [("temporal_entries", C.c_int32), ("pad", C.c_int32), ("latest_query_id", C.c_uint64),
                ("latest_island", Island)]


LCD_STATUS = ["LOOP_DETECTED", "NO_MATCHES", "LOW_NSS_FACTOR", "LOW_SCORE", "NO_GROUPS",
              "FAILED_TEMPORAL_CONSTRAINT", "FAILED_GEOM_VERIFICATION", "FAILED_POSE_RECOVERY"]

EXPORTS = [
    "kml_default_params", "kml_params_from_yaml", "kml_create", "kml_create_lane", "kml_destroy", "kml_last_error", "kml_get_stats",
    "kml_device_count", "kml_add_bow", "kml_add_bow_bulk", "kml_add_frame",
    "kml_add_frames_bulk", "kml_frame_exists", "kml_bow_exists", "kml_num_bow_for_robot",
    "kml_get_bow_vector", "kml_db_query", "kml_bow_score", "kml_detect_loop_with_robot",
    "kml_detect_loop", "kml_compute_matched_indices", "kml_geometric_verification_nister",
    "kml_recover_pose", "kml_query_batch", "kml_query_batch_upload", "kml_query_batch_run",
    "kml_host_alloc", "kml_host_free",
    "kml_hamming_knn2", "kml_l1_knn2", "kml_hamming_knn2_bench", "kml_ransac_arun_batch",
    "kml_ransac_nister_batch", "kml_ransac_onepoint_batch", "kml_vocab_set", "kml_transform_batch", "kml_peak_popc", "kml_peak_fp64", "kml_timer_begin", "kml_timer_end", "kml_flush_l2", "kml_comm_unique_id",
    "kml_comm_init", "kml_query_batch_sharded", "kml_query_batch_sharded_seq", "kml_query_batch_sharded_host", "kml_comm_seq_reset",
    "kml_merge_shard_records_device",
    "kml_save_shard", "kml_load_shard", "kml_merge_shard_records", "kml_compute_islands", "kml_check_temporal_constraint", "kml_detect_loop_islands", "kml_add_frame_msg",
]

_lib = None


def build(force=False, verbose=False):
    """Compile libkml.so in-tree with nvcc for sm_100a (works without a GPU)."""
    args = ["make", "-C", _PKG, "-s"]
    if force:
        args.append("-B")
    if verbose:
        args.append("VERBOSE=1")
    subprocess.check_call(args)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("libkml.so is not built (%s); run __graft_entry__.build() — there is "
                          "no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.kml_last_error.restype = C.c_char_p
        L.kml_last_error.argtypes = [C.c_void_p]
        _lib = L
    return _lib
