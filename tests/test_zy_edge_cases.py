"""Edge cases of the path through the C ABI against the oracle: an empty detector, databases with BoW
vectors but no frames, queries that touch no posting, a query with an empty BoW vector.  Sorted after
the established GPU tests (the file name), since it was written after the round's last GPU minute;
the CPU suite runs it against the emulated library (tests/test_emulated_library.py).
Reference behaviour restated by the oracle: DBoW2 `TemplatedDatabase::query` returns an empty
QueryResults for an empty database or a query with no shared word; `detectLoop`
returns no candidate then (A.3); verification of a candidate whose VLC frame is absent is skipped
(`frameExists`, A.3)."""
import numpy as np
import pytest

from conftest import fill
from test_gpu_parity import _check_records

pytestmark = pytest.mark.gpu


def _batch_args(q, sel=None):
    fq, fp = q["frames"], q["prev"]
    return (q["q_robot"], q["q_pose"], fq["bow_off"], fq["bow_ids"], fq["bow_vals"], fp["bow_off"],
            fp["bow_ids"], fp["bow_vals"], fq["desc"], fq["bearings"], fq["points"])


def test_empty_detector(oracle, small_world):
    import kml
    world, chunks, q = small_world
    det, lcd = kml.LoopClosureDetector(), oracle.LoopClosureDetector()
    fq = q["frames"]
    ids, vals = fq["bow_ids"][fq["bow_off"][0]:fq["bow_off"][1]], fq["bow_vals"][fq["bow_off"][0]:fq["bow_off"][1]]
    assert det.numBoWForRobot(0) == 0 and not det.bowExists(0, 0) and not det.frameExists(0, 0)
    assert det.totalBoWMatches() == 0
    ok, r1, _, _ = det.detectLoop(0, 1, ids, vals)
    r0, _, _ = lcd.detectLoop(0, 1, ids, vals)
    assert len(r0) == 0 and not ok and len(r1) == 0
    e1, sc1 = det.dbQuery(0, ids, vals, 10, -1)
    assert len(e1) == 0 and len(sc1) == 0
    out0, cnt0 = lcd.query_batch(*_batch_args(q))
    out1, cnt1 = det.query_batch(*_batch_args(q))
    assert np.array_equal(cnt0, cnt1) and not cnt1.any()
    det.close()


def test_bow_without_frames_and_disjoint_words(oracle, small_world):
    """Databases filled with BoW vectors only: candidates exist, none can be verified.  Then a batch
    whose word ids lie above every id in the databases, and a batch of empty BoW vectors."""
    import kml
    world, chunks, q = small_world
    det, lcd = kml.LoopClosureDetector(), oracle.LoopClosureDetector()
    for ch in chunks:
        for i, p in enumerate(ch["poses"][:120]):
            o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
            det.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
            lcd.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
    args = list(_batch_args(q))
    out0, cnt0 = lcd.query_batch(*args)
    out1, cnt1 = det.query_batch(*args)
    _check_records(out0, cnt0, out1, cnt1)
    assert cnt0.sum() > 0 and (out0["status"][0, :cnt0[0]] == 3).all()  # candidates, every frame missing

    # word ids no database holds: no posting is touched
    far = list(args)
    top = max(int(ch["bow_ids"].max()) for ch in chunks)
    far[3] = (np.asarray(args[3]).astype(np.int64) + top + 1).astype(np.asarray(args[3]).dtype)
    out0, cnt0 = lcd.query_batch(*far)
    out1, cnt1 = det.query_batch(*far)
    assert np.array_equal(cnt0, cnt1) and not cnt1.any()

    # empty BoW vectors for every query of the batch
    B = len(q["q_pose"])
    emp = list(args)
    emp[2] = np.zeros(B + 1, np.asarray(args[2]).dtype)
    emp[3] = np.asarray(args[3])[:0]
    emp[4] = np.asarray(args[4])[:0]
    out0, cnt0 = lcd.query_batch(*emp)
    out1, cnt1 = det.query_batch(*emp)
    assert np.array_equal(cnt0, cnt1) and not cnt1.any()
    det.close()
