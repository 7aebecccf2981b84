"""Vocabulary transform: oracle pinned by a direct numpy restatement on a
small tree (CPU), GPU kernel bit-exact against the oracle at k=10, L=6 (GPU)."""
import numpy as np
import pytest


def _np_transform(k, L, nodes, w, desc):
    popc = np.array([bin(i).count("1") for i in range(256)], np.int32)
    v = {}
    for f in desc:
        idx, off, lvl = 0, 0, 1
        for _ in range(L):
            lvl *= k
            ch = nodes[off + idx * k: off + idx * k + k]
            d = popc[ch ^ f].sum(axis=1)
            idx = idx * k + int(np.argmin(d))      # first minimum
            off += lvl
        if w[idx] > 0:
            v[idx] = v.get(idx, 0.0) + w[idx]
    ids = sorted(v)
    norm = 0.0
    for i in ids:
        norm += abs(v[i])
    return np.array(ids, np.uint32), np.array([v[i] / norm for i in ids])


def test_oracle_transform_matches_numpy(oracle):
    from kml import synth
    nodes, w = synth.make_vocabulary(4, 3, key=1)
    voc = oracle.Vocabulary(4, 3, nodes, w)
    rng = np.random.default_rng(2)
    for _ in range(5):
        desc = rng.integers(0, 256, (300, 32), np.uint8)
        desc[10] = nodes[0]                       # exact hit on a level-1 node
        ids, vals = voc.transform(desc)
        eids, evals = _np_transform(4, 3, nodes, w, desc)
        assert np.array_equal(ids, eids) and np.array_equal(vals, evals)
        assert abs(vals.sum() - 1.0) < 1e-12 and (np.diff(ids.astype(np.int64)) > 0).all()


@pytest.mark.gpu
def test_gpu_transform_bit_exact(oracle, gpu_lcd):
    from kml import synth
    nodes, w = synth.make_vocabulary(10, 6)
    voc = oracle.Vocabulary(10, 6, nodes, w)
    gpu_lcd.setVocabulary(10, 6, nodes, w)
    rng = np.random.default_rng(3)
    desc = rng.integers(0, 256, (24, 500, 32), np.uint8)
    desc[0, :40] = desc[0, 40:80]                 # repeated words inside one frame
    desc[1, :] = desc[1, 0]                       # a single word
    leaf0 = sum(10 ** l for l in range(1, 6))
    desc[2, :100] = nodes[leaf0: leaf0 + 100]     # exact leaf descriptors
    off, ids, vals, ms = gpu_lcd.transform(desc)
    for b in range(desc.shape[0]):
        eids, evals = voc.transform(desc[b])
        assert np.array_equal(ids[off[b]:off[b + 1]], eids)
        assert np.array_equal(vals[off[b]:off[b + 1]], evals)       # bit-exact
    assert off[2] - off[1] == 1
    # the transform output feeds the hot path: add it as a BoW vector and query it back
    gpu_lcd.addBowVector(77, 0, ids[off[0]:off[1]], vals[off[0]:off[1]].astype(np.float32))
    e, s = gpu_lcd.dbQuery(77, ids[off[0]:off[1]], vals[off[0]:off[1]].astype(np.float32), 1)
    assert len(e) == 1 and e[0] == 0 and abs(s[0] - 1.0) < 1e-6
