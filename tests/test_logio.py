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
