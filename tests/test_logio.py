"""loop_closures.csv written by kml.logio parses exactly the way
/root/reference/evaluation/lc_result.py:115-138 parses it (same column names,
same int()/float() conversions, inter-robot filter)."""
import csv

import numpy as np
from scipy.spatial.transform import Rotation as Rot


def test_loop_closures_csv_schema(tmp_path):
    import kml
    rec = np.zeros((2, 4), kml.RESULT_DTYPE)
    counts = np.array([2, 1], np.int32)
    R = Rot.from_rotvec([0.1, -0.2, 0.3]).as_matrix()
    T = np.c_[R, [1.0, 2.0, 3.0]]
    rec[0, 0] = (1, 10, 2, 20, 1.25, 200, 150, 90, 0, R.ravel(), T.ravel())
    rec[0, 1] = (1, 10, 1, 5, 1.0, 6, 0, 0, 1, np.zeros(9), np.zeros(12))     # rejected: not logged
    rec[1, 0] = (3, 11, 3, 400, 0.9, 210, 160, 95, 0, R.ravel(), T.ravel())   # intra-robot
    p = tmp_path / "loop_closures.csv"
    assert kml.logio.write_loop_closures_csv(str(p), rec, counts, stamps_ns=[111, 222]) == 2
    inter = []
    with open(p) as f:                      # the reference's parse_csv_files()
        for row in csv.DictReader(f):
            if row["robot1"] != row["robot2"]:
                inter.append({k: (float(row[k]) if k in ("qx", "qy", "qz", "qw", "tx", "ty", "tz", "norm_bow_score")
                                  else int(row[k])) for k in kml.logio.HEADER})
    assert len(inter) == 1
    r = inter[0]
    assert (r["robot1"], r["pose1"], r["robot2"], r["pose2"]) == (1, 10, 2, 20)
    assert (r["mono_inliers"], r["stereo_inliers"], r["stamp_ns"]) == (150, 90, 111)
    q = Rot.from_quat([r["qx"], r["qy"], r["qz"], r["qw"]]).as_matrix()
    assert np.abs(q - R).max() < 1e-12 and (r["tx"], r["ty"], r["tz"]) == (1.0, 2.0, 3.0)


REF_LC_RESULT = "/root/reference/evaluation/lc_result.py"


def _reference_parse_csv_files():
    """The reference's own parse_csv_files, taken from its source by name (the module itself
    imports matplotlib / pandas / tqdm at the top and cannot be imported here)."""
    import ast
    import os
    src = open(REF_LC_RESULT).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "parse_csv_files")
    ns = {"csv": csv, "os": os}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_LC_RESULT, "exec"), ns)
    return ns["parse_csv_files"]


def _three_files(tmp_path):
    """The three per-robot logs lc_result.py reads, written by kml.logio from one detection sequence."""
    import kml
    from kml import logio
    R1 = Rot.from_rotvec([0.1, -0.2, 0.3]).as_matrix()
    R2 = Rot.from_rotvec([-2.9, 0.4, 0.1]).as_matrix()     # qw near 0: the trace <= 0 branch
    T1, T2 = np.c_[R1, [1.0, 2.0, 3.0]], np.c_[R2, [-4.5, 0.25, 7.0]]
    rec = np.zeros((3, 2), kml.RESULT_DTYPE)
    counts = np.array([2, 1, 1], np.int32)
    rec[0, 0] = (1, 10, 2, 20, 1.25, 200, 150, 90, 0, R1.ravel(), T1.ravel())   # inter-robot, accepted
    rec[0, 1] = (1, 10, 4, 7, 1.0, 6, 0, 0, 1, np.zeros(9), np.zeros(12))      # rejected: not logged
    rec[1, 0] = (1, 11, 1, 3, 0.9, 210, 160, 95, 0, R2.ravel(), T2.ravel())    # intra-robot, accepted
    rec[2, 0] = (1, 12, 5, 40, 0.7, 180, 33, 21, 0, R2.ravel(), T2.ravel())    # inter-robot, accepted
    d = tmp_path / "robot1" / "distributed"
    d.mkdir(parents=True)
    lc, st, rs = str(d / "loop_closures.csv"), str(d / "output_lcd_status.csv"), str(d / "output_lcd_result.csv")
    assert logio.write_loop_closures_csv(lc, rec, counts, stamps_ns=[111, 222, 333]) == 3
    status = [{"timestamp_kf": 100, "lcd_status": "NO_MATCHES", "query_id": 9, "match_id": 0},
              {"timestamp_kf": 110, "lcd_status": "LOOP_DETECTED", "query_id": 11, "match_id": 3,
               "mono_inliers": 160, "stereo_inliers": 95},
              {"timestamp_kf": 120, "lcd_status": "FAILED_GEOM_VERIFICATION", "query_id": 12, "match_id": 4,
               "mono_inliers": 3},
              {"timestamp_kf": 130, "lcd_status": "FAILED_POSE_RECOVERY", "query_id": 13, "match_id": 5,
               "mono_inliers": 40, "stereo_inliers": 2},
              {"timestamp_kf": 140, "lcd_status": "LOOP_DETECTED", "query_id": 14, "match_id": 6,
               "mono_inliers": 70, "stereo_inliers": 31}]
    assert all(s["lcd_status"] in kml.LCD_STATUS for s in status)
    assert logio.write_lcd_status_csv(st, status) == 5
    result = [{"timestamp_kf": 100, "timestamp_query": 100, "timestamp_match": 0, "isLoop": 0, "matchKfId": 0, "queryKfId": 9},
              {"timestamp_kf": 110, "timestamp_query": 110, "timestamp_match": 30, "isLoop": 1, "matchKfId": 3,
               "queryKfId": 11, "T": T2.ravel()},
              {"timestamp_kf": 120, "timestamp_query": 120, "timestamp_match": 40, "isLoop": 0, "matchKfId": 4, "queryKfId": 12},
              {"timestamp_kf": 130, "timestamp_query": 130, "timestamp_match": 50, "isLoop": 0, "matchKfId": 5, "queryKfId": 13},
              {"timestamp_kf": 140, "timestamp_query": 140, "timestamp_match": 60, "isLoop": 1, "matchKfId": 6,
               "queryKfId": 14, "x": 0.5, "y": -1.5, "z": 2.5, "qx": 0.0, "qy": 0.0, "qz": 0.0, "qw": 1.0}]
    assert logio.write_lcd_result_csv(rs, result) == 5
    return (lc, st, rs), (R1, R2, T1, T2)


def _check_parsed(inter, intra, rejected, mats):
    R1, R2, T1, T2 = mats
    assert [(r["robot1"], r["pose1"], r["robot2"], r["pose2"], r["stamp_ns"]) for r in inter] == \
        [(1, 10, 2, 20, 111), (1, 12, 5, 40, 333)]
    assert [(r["mono_inliers"], r["stereo_inliers"]) for r in inter] == [(150, 90), (33, 21)]
    assert inter[0]["norm_bow_score"] == 1.25
    for r, R, T in ((inter[0], R1, T1), (inter[1], R2, T2)):
        q = Rot.from_quat([r["qx"], r["qy"], r["qz"], r["qw"]]).as_matrix()
        assert np.abs(q - R).max() < 1e-12 and (r["tx"], r["ty"], r["tz"]) == tuple(T[:, 3])
    assert [(r["pose2"], r["pose1"], r["mono_inliers"], r["stereo_inliers"]) for r in intra] == \
        [(11, 3, 160, 95), (14, 6, 70, 31)]
    assert [(r["timestamp2"], r["timestamp1"]) for r in intra] == [(110, 30), (140, 60)]
    q = Rot.from_quat([intra[0][k] for k in ("qx", "qy", "qz", "qw")]).as_matrix()
    assert np.abs(q - R2).max() < 1e-12
    assert (intra[0]["tx"], intra[0]["ty"], intra[0]["tz"]) == tuple(T2[:, 3])
    assert (intra[1]["tx"], intra[1]["ty"], intra[1]["tz"], intra[1]["qw"]) == (0.5, -1.5, 2.5, 1.0)
    assert [(r["pose2"], r["pose1"], r["lcd_status"], r["mono_inliers"], r["stereo_inliers"]) for r in rejected] == \
        [(12, 4, "FAILED_GEOM_VERIFICATION", 3, 0), (13, 5, "FAILED_POSE_RECOVERY", 40, 2)]


def test_three_logs_parse_like_lc_result(tmp_path):
    """loop_closures.csv + output_lcd_status.csv + output_lcd_result.csv, read back with a
    restatement of /root/reference/evaluation/lc_result.py:115-183 (parse_csv_files): the
    inter-robot filter, the LOOP_DETECTED / FAILED_* split and the in-order pairing of isLoop rows
    with LOOP_DETECTED rows (the reference asserts on it)."""
    (lc, st, rs), mats = _three_files(tmp_path)
    inter, intra, rejected = [], [], []
    with open(lc) as f:
        for row in csv.DictReader(f):
            if row["robot1"] != row["robot2"]:
                inter.append({k: (float(row[k]) if k in ("qx", "qy", "qz", "qw", "tx", "ty", "tz", "norm_bow_score")
                                  else int(row[k])) for k in row})
    with open(st) as f:
        for row in csv.DictReader(f):
            rec = {"pose2": int(row["query_id"]), "pose1": int(row["match_id"]),
                   "mono_inliers": int(row["mono_inliers"]), "stereo_inliers": int(row["stereo_inliers"])}
            if row["lcd_status"] == "LOOP_DETECTED":
                intra.append(rec)
            elif row["lcd_status"] in ("FAILED_TEMPORAL_CONSTRAINT", "FAILED_GEOM_VERIFICATION", "FAILED_POSE_RECOVERY"):
                rejected.append(dict(rec, lcd_status=row["lcd_status"]))
    with open(rs) as f:
        k = 0
        for row in csv.DictReader(f):
            if row["isLoop"] == "1":
                assert (intra[k]["pose2"], intra[k]["pose1"]) == (int(row["queryKfId"]), int(row["matchKfId"]))
                intra[k].update(timestamp2=int(row["timestamp_query"]), timestamp1=int(row["timestamp_match"]),
                                tx=float(row["x"]), ty=float(row["y"]), tz=float(row["z"]), qx=float(row["qx"]),
                                qy=float(row["qy"]), qz=float(row["qz"]), qw=float(row["qw"]))
                k += 1
    assert k == len(intra)
    _check_parsed(inter, intra, rejected, mats)


def test_three_logs_through_the_reference_parser(tmp_path):
    """The same three files through the reference's own parse_csv_files, unchanged (only where the
    reference tree is present: this container, not the GPU box)."""
    import os
    import pytest
    if not os.path.exists(REF_LC_RESULT):
        pytest.skip("reference tree not present")
    (lc, st, rs), mats = _three_files(tmp_path)
    inter, intra, rejected = _reference_parse_csv_files()(lc, st, rs)
    _check_parsed(inter, intra, rejected, mats)
