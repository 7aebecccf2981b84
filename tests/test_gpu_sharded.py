"""The robot-sharded query on real GPUs (needs >= 2 B200s; skipped on a single-GPU box): two
processes, one GPU and one robot-database shard each, the same query batch on both;
kml_query_batch_sharded (in-place ncclAllGather of the record blocks + merge_shards_kernel) must
return, on every rank, the merge of both ranks' local records by the documented rule
(kml/shard.py) and the oracle's records over the union of the databases (SURVEY.md §8e:
sharding by robot is exact).  Also: the device merge kernel alone against the Python merge rule
(runs on one GPU and under the CPU emulator)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_merge_kernel_follows_the_merge_rule(gpu_lcd):
    import kml
    from kml import shard
    rng = np.random.default_rng(5)
    for nranks, B, cap_in, cap in [(2, 4, 3, 4), (8, 16, 6, 10), (8, 33, 16, 16), (3, 5, 16, 40)]:
        blocks = []
        for r in range(nranks):
            rec = np.zeros((B, cap_in), kml.RESULT_DTYPE)
            cnt = rng.integers(0, cap_in + 1, B).astype(np.int32)
            rec["norm_bow_score"] = rng.integers(0, 5, (B, cap_in)) / 4.0     # many exact ties
            rec["m_robot"] = r if r != 1 else 0                              # ranks 0 and 1 hold the same robot: key ties
            rec["m_pose"] = rng.integers(0, 12, (B, cap_in))
            rec["mono_inliers"] = rng.integers(0, 99, (B, cap_in))
            rec["T"] = rng.normal(size=(B, cap_in, 12))
            for b in range(B):   # per-rank lists arrive ranked, as the library emits them
                o = np.lexsort((rec["m_pose"][b, :cnt[b]], rec["m_robot"][b, :cnt[b]], -rec["norm_bow_score"][b, :cnt[b]]))
                rec[b, :cnt[b]] = rec[b, :cnt[b]][o]
            blocks.append((rec, cnt))
        o1, c1 = shard.merge_records(blocks, cap)
        o2, c2 = gpu_lcd.merge_shard_records_device(blocks, cap)
        assert np.array_equal(c1, c2)
        for b in range(B):
            assert o1[b, :c1[b]].tobytes() == o2[b, :c2[b]].tobytes(), (nranks, b)
            assert not o2[b, c2[b]:].tobytes().strip(b"\0")               # slots beyond the count stay zero


def _rank(rank, world, port, key, q):
    sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kml
    from kml import shard, synth
    from kml.rendezvous import Rendezvous
    try:
        r = Rendezvous(rank=rank, world=world, addr="127.0.0.1", port=port, key=key)
        prm = kml.default_params()
        prm.top_k_verify = 6
        det = kml.LoopClosureDetector(prm, device=rank)
        lane = det.create_lane()
        uids = r.broadcast([kml.LoopClosureDetector.comm_unique_id() for _ in range(2)] if rank == 0 else None)
        det.comm_init(world, rank, uids[0])
        lane.comm_init(world, rank, uids[1])
        wd = synth.World(60, F=500)
        for ch in synth.build_database(wd, shard.robots_of_rank(rank, 1), 240, chunk=240):
            det.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
            det.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
        ok, n_rec = True, 0
        for k, handle in enumerate((det, lane, det)):
            qs = synth.make_queries(wd, 24, 240, world, key=k)          # identical on every rank
            fq, fp = qs["frames"], qs["prev"]
            args = (qs["q_robot"], qs["q_pose"], fq["bow_off"], fq["bow_ids"], fq["bow_vals"], fp["bow_off"],
                    fp["bow_ids"], fp["bow_vals"], fq["desc"], fq["bearings"], fq["points"])
            handle.query_batch_upload(*args)
            loc, lcnt = handle.query_batch_run(sharded=False)
            if k == 2:
                mrg, mcnt = handle.query_batch(*args, sharded=True)       # host buffers in, un-sequenced
            else:
                mrg, mcnt = handle.query_batch_run(sharded=True, seq=k)   # sequenced, one batch per lane
            blocks = r.gather((loc, lcnt))
            ref, rcnt = shard.merge_records(blocks, int(prm.top_k_verify))
            same = np.array_equal(rcnt, mcnt) and all(ref[b, :rcnt[b]].tobytes() == mrg[b, :mcnt[b]].tobytes()
                                                      for b in range(len(rcnt)))
            ok = ok and bool(same)
            n_rec += int(mcnt.sum())
            if rank == 0 and k == 0:                                     # and the oracle over both databases
                import kml_oracle as ko
                kml_oracle_prm = ko.default_params()
                kml_oracle_prm.top_k_verify = 6
                full = ko.LoopClosureDetector(kml_oracle_prm)
                for ch in synth.build_database(wd, range(world), 240, chunk=240):
                    for i, p in enumerate(ch["poses"]):
                        o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
                        full.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
                        full.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
                o0_, c0_ = full.query_batch(*args, threads=4)
                ok = ok and np.array_equal(c0_, mcnt)
                for b in range(len(c0_)):
                    for i in range(c0_[b]):
                        for f in ("m_robot", "m_pose", "n_matches", "mono_inliers", "stereo_inliers", "status"):
                            ok = ok and o0_[b, i][f] == mrg[b, i][f]
        r.barrier()
        det.close()
        r.close()
        q.put((rank, bool(ok), n_rec))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))


def test_two_gpu_sharded_query_equals_merge_of_local_records():
    import kml
    if kml.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    key = "shard_%d" % os.getpid()
    procs = [ctx.Process(target=_rank, args=(r, 2, 29211, key, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] and res[1][1], res
    assert res[0][2] > 20 and res[0][2] == res[1][2]
