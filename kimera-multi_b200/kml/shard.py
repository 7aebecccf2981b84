"""Robot-sharded query plumbing shared by bench.py and the tests.

Ownership rule: robot r lives on rank r // robots_per_rank.  Every rank runs the
same query batch against its own shard; the per-rank record blocks are merged
per query by (normalised score desc, match robot asc, match pose asc) and cut to
`cap` — the rule kml_query_batch_sharded applies after its ncclAllGather
(kimera-multi_b200/csrc/lcd.cu).
"""
import numpy as np


def owner_rank(robot, robots_per_rank):
    return int(robot) // int(robots_per_rank)


def robots_of_rank(rank, robots_per_rank):
    return [rank * robots_per_rank + r for r in range(robots_per_rank)]


def merge_records(blocks, cap):
    """blocks: list over ranks of (records[B, cap_r], counts[B]) -> merged (records[B, cap], counts[B])."""
    B = len(blocks[0][1])
    out = np.zeros((B, cap), blocks[0][0].dtype)
    counts = np.zeros(B, np.int32)
    for b in range(B):
        rows = [rec[b, i] for rec, cnt in blocks for i in range(int(cnt[b]))]
        rows.sort(key=lambda r: (-float(r["norm_bow_score"]), int(r["m_robot"]), int(r["m_pose"])))
        n = min(len(rows), cap)
        for i in range(n):
            out[b, i] = rows[i]
        counts[b] = n
    return out, counts


def merge_records_native(blocks, cap):
    """The same merge through libkml.so (kml_merge_shard_records) — what kml_query_batch_sharded
    applies to the all-gathered blocks."""
    import ctypes as C
    from . import _lib
    B = len(blocks[0][1])
    cap_in = blocks[0][0].shape[1]
    buf = b"".join(np.ascontiguousarray(rec).tobytes() + np.ascontiguousarray(cnt, np.int32).tobytes()
                   for rec, cnt in blocks)
    out = np.zeros((B, cap), blocks[0][0].dtype)
    counts = np.zeros(B, np.int32)
    rc = _lib.lib().kml_merge_shard_records(buf, len(blocks), B, cap_in, cap,
                                           out.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("kml_merge_shard_records: %d" % rc)
    return out, counts
