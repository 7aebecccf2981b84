"""GPU parity for a robot database wider than one accumulator tile of the BoW scorer
(BASELINE.json configs[4] holds 50 000 keyframes per database; bow.cu keeps 24 576 entries per
CTA).  Written after this round's GPU minutes were spent: the oracle half was run on the CPU, the
GPU half has not run yet, which is why the file sorts after the established tests under -x."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_database_spanning_several_entry_tiles(oracle):
    """One robot database of 60 000 entries (the BoW scorer keeps 24 576 accumulators per CTA, so
    three entry tiles; C5 holds 50 000 keyframes per database): Database::query against the oracle,
    with exact duplicates of one vector placed in different tiles (equal scores -> ascending entry
    id across the tile merge) and max_id cuts inside the first and the last tile."""
    import kml
    from test_gpu_parity import _same_order_modulo_ties, BOW_RTOL
    rng = np.random.default_rng(77)
    n, words, vocab = 60000, 40, 3000
    # distinct ascending word ids per entry without an n x vocab matrix: sorted sample of a stride grid
    base = rng.integers(0, vocab // words, (n, words))
    ids = (np.arange(words)[None, :] * (vocab // words) + base).astype(np.uint32)
    vals = rng.random((n, words)).astype(np.float32) + np.float32(0.01)
    vals = (vals / vals.sum(axis=1, keepdims=True)).astype(np.float32)
    dup = [5, 24575, 24576, 49151, 49152, 59999]          # both sides of each tile boundary
    for e in dup[1:]:
        ids[e], vals[e] = ids[dup[0]], vals[dup[0]]
    off = (np.arange(n + 1) * words).astype(np.int64)
    det = kml.LoopClosureDetector()
    det.addBowVectors(3, np.arange(n, dtype=np.uint64), off, ids.reshape(-1), vals.reshape(-1))
    db = oracle.Database()
    for i in range(n):
        db.add(ids[i], vals[i])
    queries = [(ids[dup[0]], vals[dup[0]])]
    for k in range(6):
        qi = ids[1000 + 9000 * k].copy()
        qv = rng.random(words).astype(np.float32)
        queries.append((qi, (qv / qv.sum()).astype(np.float32)))
    for qi, qv in queries:
        for max_results, max_id in [(50, -1), (128, -1), (1, -1), (50, 30000), (50, 100), (50, 59000)]:
            e0, s0 = db.query(qi, qv, max_results, max_id)
            e1, s1 = det.dbQuery(3, qi, qv, max_results, max_id)
            assert len(e0) == len(e1) and len(e0) > 0
            np.testing.assert_allclose(s1, s0, rtol=BOW_RTOL, atol=0)
            assert _same_order_modulo_ties(e0, s0, e1, s1)
    # the duplicates score 1 against their own vector and come back in ascending entry id
    e1, s1 = det.dbQuery(3, queries[0][0], queries[0][1], 50, -1)
    assert list(e1[:len(dup)]) == dup and np.all(np.abs(s1[:len(dup)] - 1.0) < 1e-6)
    det.close()
