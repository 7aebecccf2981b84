"""Evaluation-compatible logging.

Writes the `loop_closures.csv` the front end logs per robot, with exactly the
header /root/reference/evaluation/lc_result.py:115-138 reads through
csv.DictReader: robot1,pose1,robot2,pose2,qx,qy,qz,qw,tx,ty,tz,norm_bow_score,
mono_inliers,stereo_inliers,stamp_ns — so lc_result.py / analyze_inliers.py run
unchanged on records produced by libkml.so.
"""
import csv

import numpy as np

HEADER = ["robot1", "pose1", "robot2", "pose2", "qx", "qy", "qz", "qw", "tx", "ty", "tz",
          "norm_bow_score", "mono_inliers", "stereo_inliers", "stamp_ns"]


def rot_to_quat(R):
    """Row-major 3x3 rotation -> (qx, qy, qz, qw), qw >= 0 (gtsam::Rot3::toQuaternion order)."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = 2.0 * np.sqrt(1.0 + tr)
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = 2.0 * np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def write_loop_closures_csv(path, records, counts, stamps_ns=None, append=False):
    """One row per verified loop closure (status == 0) of a kml.RESULT_DTYPE batch.
    robot1/pose1 = query keyframe, robot2/pose2 = match keyframe, pose = T_query_match."""
    n = 0
    with open(path, "a" if append else "w", newline="") as f:
        w = csv.writer(f)
        if not append:
            w.writerow(HEADER)
        for b in range(len(counts)):
            for i in range(int(counts[b])):
                r = records[b, i]
                if r["status"] != 0:
                    continue
                T = np.asarray(r["T"]).reshape(3, 4)
                q = rot_to_quat(T[:, :3])
                stamp = 0 if stamps_ns is None else int(stamps_ns[b])
                w.writerow([int(r["q_robot"]), int(r["q_pose"]), int(r["m_robot"]), int(r["m_pose"]),
                            repr(float(q[0])), repr(float(q[1])), repr(float(q[2])), repr(float(q[3])),
                            repr(float(T[0, 3])), repr(float(T[1, 3])), repr(float(T[2, 3])),
                            repr(float(r["norm_bow_score"])), int(r["mono_inliers"]),
                            int(r["stereo_inliers"]), stamp])
                n += 1
    return n


STATUS_HEADER = ["timestamp_kf", "lcd_status", "query_id", "match_id", "mono_input_size", "mono_inliers",
                 "mono_iters", "stereo_input_size", "stereo_inliers", "stereo_iters"]


def write_lcd_status_csv(path, rows, append=False):
    """`output_lcd_status.csv` of the intra-robot detector, read by
    /root/reference/evaluation/lc_result.py:143-162 through csv.DictReader (columns used:
    lcd_status, query_id, match_id, mono_inliers, stereo_inliers).  rows = iterable of dicts with
    those keys; lcd_status is one of kml.LCD_STATUS; missing keys are written as 0."""
    n = 0
    with open(path, "a" if append else "w", newline="") as f:
        w = csv.writer(f)
        if not append:
            w.writerow(STATUS_HEADER)
        for r in rows:
            w.writerow([r.get(k, 0) for k in STATUS_HEADER])
            n += 1
    return n


RESULT_HEADER = ["timestamp_kf", "timestamp_query", "timestamp_match", "isLoop", "matchKfId", "queryKfId",
                 "x", "y", "z", "qw", "qx", "qy", "qz"]


def write_lcd_result_csv(path, rows, append=False):
    """`output_lcd_result.csv` of the intra-robot detector: one row per processed keyframe, the
    third file /root/reference/evaluation/lc_result.py:165-183 reads (columns used: isLoop,
    queryKfId, matchKfId, timestamp_query, timestamp_match, x, y, z, qx, qy, qz, qw).  That reader
    walks the isLoop == 1 rows in file order and asserts that the k-th of them names the same
    (query, match) pair as the k-th LOOP_DETECTED row of output_lcd_status.csv, so both files must
    be written from the same sequence of detections.  rows = iterable of dicts: timestamp_kf,
    timestamp_query, timestamp_match, isLoop, matchKfId, queryKfId and either T (3x4 row-major
    T_match_query, as kml.RESULT_DTYPE holds it) or x, y, z, qx, qy, qz, qw; a row without a loop
    (isLoop 0) is written with the identity pose."""
    n = 0
    with open(path, "a" if append else "w", newline="") as f:
        w = csv.writer(f)
        if not append:
            w.writerow(RESULT_HEADER)
        for r in rows:
            r = dict(r)
            if "T" in r and r["T"] is not None:
                T = np.asarray(r["T"], dtype=np.float64).reshape(3, 4)
                q = rot_to_quat(T[:, :3])
                r.update(x=float(T[0, 3]), y=float(T[1, 3]), z=float(T[2, 3]),
                         qx=float(q[0]), qy=float(q[1]), qz=float(q[2]), qw=float(q[3]))
            elif not int(r.get("isLoop", 0)):
                for k, v in (("x", 0.0), ("y", 0.0), ("z", 0.0), ("qx", 0.0), ("qy", 0.0), ("qz", 0.0), ("qw", 1.0)):
                    r.setdefault(k, v)
            out = []
            for k in RESULT_HEADER:
                v = r.get(k, 0)
                out.append(repr(float(v)) if k in ("x", "y", "z", "qw", "qx", "qy", "qz") else int(v))
            w.writerow(out)
            n += 1
    return n
