"""Row f3: island grouping, temporal-consistency check, VLCFrameMsg layout and
the status CSV — product host logic (libkml.so, no kernels) against the oracle's restatement of
Kimera-VIO's LcdThirdPartyWrapper and against hand-worked cases."""
import csv

import numpy as np
import pytest


def test_islands_hand_cases():
    import kml
    # ids 10,11,12 | 20 | 30,32 with gap 3: three islands; sums, best entries, ascending start
    ids = [30, 10, 20, 12, 11, 32]
    sc = [0.5, 0.1, 0.9, 0.3, 0.3, 0.6]
    isl = kml.compute_islands(ids, sc, 3, 1)
    assert [(a, b, c) for a, b, c, _, _ in isl] == [(10, 12, 11), (20, 20, 20), (30, 32, 32)]
    assert isl[0][3] == pytest.approx(0.7) and isl[0][4] == 0.3   # first of the tied best scores wins
    assert isl[2][3] == pytest.approx(1.1) and isl[2][4] == 0.6
    # gap is strict: ids 3 apart do not join at max_intraisland_gap = 3, they do at 4
    assert len(kml.compute_islands([1, 4], [1.0, 1.0], 3, 1)) == 2
    assert len(kml.compute_islands([1, 4], [1.0, 1.0], 4, 1)) == 1
    # length threshold counts ids spanned (end - start + 1), not results
    assert len(kml.compute_islands([1, 3, 10], [1.0, 1.0, 1.0], 3, 3)) == 1
    # a single result is always an island; no results, no islands
    assert kml.compute_islands([7], [0.2], 3, 5) == [(7, 7, 7, 0.2, 0.2)]
    assert kml.compute_islands([], [], 3, 1) == []


def test_islands_and_temporal_match_oracle(oracle):
    import kml
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(0, 50))
        ids = rng.choice(400, n, replace=False).astype(np.uint64)
        sc = rng.random(n)
        if n > 4:
            sc[rng.integers(0, n, 3)] = 0.5          # exact ties
        gap, mlen = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        a = kml.compute_islands(ids, sc, gap, mlen)
        b = oracle.compute_islands(ids, sc, gap, mlen)
        assert [x[:3] for x in a] == [x[:3] for x in b]
        # the product accumulates the island score while walking, the oracle sums afterwards: same order
        assert [x[3:] for x in a] == [x[3:] for x in b]
    # temporal constraint: random island sequences with query-id jumps
    for trial in range(50):
        sa, sb = kml.TemporalState(), oracle.TemporalState()
        qid = 10
        mq, mi, mt = int(rng.integers(1, 4)), int(rng.integers(0, 5)), int(rng.integers(0, 4))
        for step in range(40):
            qid += int(rng.integers(1, 5))
            s0 = int(rng.integers(0, 60)); e0 = s0 + int(rng.integers(0, 6))
            isl = (s0, e0, s0, 1.0, 1.0)
            ra = kml.check_temporal_constraint(sa, qid, isl, mq, mi, mt)
            rb = oracle.check_temporal_constraint(sb, qid, isl, mq, mi, mt)
            assert ra == rb and sa.temporal_entries == sb.temporal_entries


def test_temporal_hand_case():
    import kml
    st = kml.TemporalState()
    # min_temporal_matches = 1: the first consistent PAIR of queries passes
    assert not kml.check_temporal_constraint(st, 100, (10, 12, 11, 1.0, 0.5), 2, 3, 1)
    assert kml.check_temporal_constraint(st, 101, (12, 14, 13, 1.0, 0.5), 2, 3, 1)      # overlap
    assert kml.check_temporal_constraint(st, 103, (17, 18, 17, 1.0, 0.5), 2, 3, 1)      # gap 3 <= 3
    assert not kml.check_temporal_constraint(st, 104, (40, 41, 40, 1.0, 0.5), 2, 3, 1)  # far island: reset
    assert not kml.check_temporal_constraint(st, 110, (41, 42, 41, 1.0, 0.5), 2, 3, 1)  # query gap 6 > 2: reset


def test_status_csv_is_readable_by_the_reference_reader(tmp_path):
    from kml import logio
    import kml
    rows = [{"lcd_status": kml.LCD_STATUS[0], "query_id": 120, "match_id": 17, "mono_inliers": 80, "stereo_inliers": 55},
            {"lcd_status": kml.LCD_STATUS[5], "query_id": 121, "match_id": 18},
            {"lcd_status": kml.LCD_STATUS[6], "query_id": 122, "match_id": 19, "mono_inliers": 3}]
    path = tmp_path / "output_lcd_status.csv"
    assert logio.write_lcd_status_csv(str(path), rows) == 3
    # the loop of /root/reference/evaluation/lc_result.py:143-162
    ok, rejected = [], []
    with open(path) as f:
        for row in csv.DictReader(f):
            if row["lcd_status"] == "LOOP_DETECTED":
                ok.append((int(row["query_id"]), int(row["match_id"]), int(row["mono_inliers"]), int(row["stereo_inliers"])))
            elif row["lcd_status"] in ("FAILED_TEMPORAL_CONSTRAINT", "FAILED_GEOM_VERIFICATION", "FAILED_POSE_RECOVERY"):
                rejected.append((int(row["query_id"]), int(row["match_id"]), row["lcd_status"], int(row["mono_inliers"])))
    assert ok == [(120, 17, 80, 55)]
    assert rejected == [(121, 18, "FAILED_TEMPORAL_CONSTRAINT", 0), (122, 19, "FAILED_GEOM_VERIFICATION", 3)]


@pytest.mark.gpu
def test_detect_loop_islands_and_frame_msg(oracle, oracle_lcd, gpu_lcd, small_world):
    """kml_detect_loop_islands == oracle detectLoopWithRobot + oracle islands + oracle temporal check."""
    import kml
    world, chunks, queries = small_world
    fq, fp = queries["frames"], queries["prev"]

    def bow(fr, i):
        o0, o1 = fr["bow_off"][i], fr["bow_off"][i + 1]
        return fr["bow_ids"][o0:o1], fr["bow_vals"][o0:o1]

    n_detected = 0
    for robot in (0, 1):
        sa, sb = kml.TemporalState(), oracle.TemporalState()
        for b in range(len(queries["q_pose"])):
            qr, qp = int(queries["q_robot"][b]), int(queries["q_pose"][b])
            pids, pvals = bow(fp, b)
            oracle_lcd.addBowVector(qr, qp - 1, pids, pvals)
            gpu_lcd.addBowVector(qr, qp - 1, pids, pvals)
            ids, vals = bow(fq, b)
            status, mp, ms, isl = gpu_lcd.detectLoopIslands(robot, qr, qp, ids, vals, sa, 3, 1, 3, 0)
            r0, p0, s0 = oracle_lcd.detectLoopWithRobot(robot, qr, qp, ids, vals)
            if len(p0) == 0:
                assert status in ("NO_MATCHES", "LOW_NSS_FACTOR", "LOW_SCORE")
                continue
            oi = oracle.compute_islands(p0, s0, 3, 1)
            best = max(range(len(oi)), key=lambda i: (oi[i][3], -i))
            passed = oracle.check_temporal_constraint(sb, qp, oi[best], gpu_lcd.params.max_nrFrames_between_queries, 3, 0)
            assert status == ("LOOP_DETECTED" if passed else "FAILED_TEMPORAL_CONSTRAINT")
            assert mp == oi[best][2] and isl[:3] == oi[best][:3]
            assert abs(ms - oi[best][4]) <= 1e-6 * abs(oi[best][4])
            n_detected += passed
    assert n_detected > 0
    # VLCFrameMsg layout: float32 clouds widen to the same stored frame as their double images
    ch = chunks[0]
    d, v, k = ch["desc"][3], ch["bearings"][3].astype(np.float32), ch["points"][3].astype(np.float32)
    gpu_lcd.addVLCFrameMsg(9, 1, d, v, k)
    gpu_lcd.addVLCFrame(9, 2, d, v.astype(np.float64), k.astype(np.float64))
    iq, im = gpu_lcd.computeMatchedIndices(9, 1, 9, 2)
    assert len(iq) > 0 and np.array_equal(iq, im)           # identical descriptors match one to one
    ok, jq, jm, T = gpu_lcd.recoverPose(9, 1, 9, 2, iq, im)
    assert ok and np.abs(T[:, :3] - np.eye(3)).max() < 1e-9 and np.abs(T[:, 3]).max() < 1e-9
