"""GPU parity for a robot database wider than one accumulator tile of the BoW scorer
(BASELINE.json configs[4] holds 50 000 keyframes per database; bow.cu keeps at most 20 480 entries
per CTA), and for the incremental inverted file: vectors appended in place between queries."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_database_spanning_several_entry_tiles(oracle):
    """One robot database of 60 000 entries (the BoW scorer keeps at most 20 480 accumulators per CTA, so
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
    T = (n + 2) // 3
    T = (T + 255) // 256 * 256                            # bow_tiling: three equal tiles, a multiple of 256
    dup = [5, T - 1, T, 2 * T - 1, 2 * T, 59999]          # both sides of each tile boundary
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


def test_incremental_inverted_file_interleaved_add_and_query(oracle):
    """The deployment pattern (/root/reference/launch/kimera_vio_jackal.launch:13-15: one BoW vector
    per keyframe, then a query): after a bulk load every addBowVector is appended to the resident
    inverted file in place — rows that are full move to the end of the pool — and every query must
    equal the oracle's Database::query over everything added so far.  Word ids repeat heavily
    (vocabulary of 600 words), so rows grow, relocate several times and a garbage-triggered rebuild
    happens along the way."""
    import kml
    from test_gpu_parity import _same_order_modulo_ties, BOW_RTOL
    rng = np.random.default_rng(123)
    vocab, words = 600, 30

    def vec():
        ids = np.sort(rng.choice(vocab, words, replace=False)).astype(np.uint32)
        v = rng.random(words).astype(np.float32) + np.float32(0.01)
        return ids, (v / v.sum()).astype(np.float32)

    det = kml.LoopClosureDetector()
    db = oracle.Database()
    n0 = 300
    bulk = [vec() for _ in range(n0)]
    off = (np.arange(n0 + 1) * words).astype(np.int64)
    det.addBowVectors(1, np.arange(n0, dtype=np.uint64), off, np.concatenate([b[0] for b in bulk]),
                      np.concatenate([b[1] for b in bulk]))
    for ids, vals in bulk:
        db.add(ids, vals)
    n = n0
    for step in range(260):
        k = 1 if step % 7 else 5                      # mostly add-one / query-one, sometimes a few adds per query
        for _ in range(k):
            ids, vals = vec()
            if step == 100:                           # a word beyond the row table forces one rebuild
                ids = ids.copy(); ids[-1] = 70000
            det.addBowVector(1, n, ids, vals)
            db.add(ids, vals)
            n += 1
        qi, qv = vec()
        for max_results, max_id in [(50, -1), (5, n - 3)]:
            e0, s0 = db.query(qi, qv, max_results, max_id)
            e1, s1 = det.dbQuery(1, qi, qv, max_results, max_id)
            assert len(e0) == len(e1) and len(e0) > 0, step
            np.testing.assert_allclose(s1, s0, rtol=BOW_RTOL, atol=0)
            assert _same_order_modulo_ties(e0, s0, e1, s1), step
        if step % 50 == 0:                             # the entry just added is found with score 1
            e1, s1 = det.dbQuery(1, ids, vals, 1, -1)
            assert abs(s1[0] - 1.0) < 1e-6
    assert det.numBoWForRobot(1) == n
    det.close()
