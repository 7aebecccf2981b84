"""World-size-2 (gloo, CPU) test of the robot-sharded query: two ranks each own
one robot database, run the same replicated query batch against their shard,
all-gather their record blocks and merge — the result must equal the
single-process run over both databases (sharding by robot is exact).
The per-shard work is done by the CPU oracle here; the GPU library applies the
same merge rule after its ncclAllGather."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import kml_oracle as ko
    from kml import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    world_ = synth.World(40, F=500)
    queries = synth.make_queries(world_, 4, 120, world)          # identical on every rank (same seeds)
    fq, fp = queries["frames"], queries["prev"]
    args = (queries["q_robot"], queries["q_pose"], fq["bow_off"], fq["bow_ids"], fq["bow_vals"],
            fp["bow_off"], fp["bow_ids"], fp["bow_vals"], fq["desc"], fq["bearings"], fq["points"])
    prm = ko.default_params()
    prm.top_k_verify = 6

    def build(robots):
        lcd = ko.LoopClosureDetector(prm)
        for ch in synth.build_database(world_, robots, 120, chunk=120):
            for i, p in enumerate(ch["poses"]):
                o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
                lcd.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
                lcd.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
        return lcd

    mine = build(shard.robots_of_rank(rank, 1))
    rec, cnt = mine.query_batch(*args, threads=2)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rec, cnt))
    merged, mcnt = shard.merge_records(gathered, int(prm.top_k_verify))
    uid = [b"x" * 128 if rank == 0 else None]                       # the unique-id broadcast bench.py does
    dist.broadcast_object_list(uid, src=0)
    assert uid[0] == b"x" * 128
    if rank == 0:
        full = build(list(range(world)))
        ref, rcnt = full.query_batch(*args, threads=2)
        ok = bool(np.array_equal(rcnt, mcnt))
        for b in range(len(rcnt)):
            ok = ok and ref[b, :rcnt[b]].tobytes() == merged[b, :mcnt[b]].tobytes()
        q.put((ok, int(mcnt.sum()), int((merged["status"][mcnt[:, None] > np.arange(merged.shape[1])[None]] == 0).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_merge_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n, n_lc = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n > 0 and n_lc > 0


def test_merge_rule_orders_and_caps():
    sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
    import kml
    from kml import shard
    a = np.zeros((1, 3), kml.RESULT_DTYPE)
    b = np.zeros((1, 3), kml.RESULT_DTYPE)
    a["norm_bow_score"][0, :2] = [0.9, 0.5]; a["m_robot"][0, :2] = [0, 0]; a["m_pose"][0, :2] = [7, 3]
    b["norm_bow_score"][0, :3] = [0.9, 0.7, 0.1]; b["m_robot"][0, :3] = [1, 1, 1]; b["m_pose"][0, :3] = [2, 9, 4]
    out, cnt = shard.merge_records([(a, np.array([2])), (b, np.array([3]))], 4)
    assert cnt[0] == 4
    assert out["norm_bow_score"][0].tolist() == [0.9, 0.9, 0.7, 0.5]
    assert out["m_robot"][0].tolist() == [0, 1, 1, 0]          # exact score tie -> lower robot id first
    assert shard.owner_rank(13, 6) == 2 and shard.robots_of_rank(1, 6) == [6, 7, 8, 9, 10, 11]
    # the library's merge (what kml_query_batch_sharded applies after the all-gather) follows the same rule
    out2, cnt2 = shard.merge_records_native([(a, np.array([2])), (b, np.array([3]))], 4)
    assert np.array_equal(cnt, cnt2) and out.tobytes() == out2.tobytes()
    rng = np.random.default_rng(5)
    blocks = []
    for r in range(8):
        rec = np.zeros((16, 6), kml.RESULT_DTYPE)
        cntr = rng.integers(0, 7, 16).astype(np.int32)
        rec["norm_bow_score"] = rng.integers(0, 5, (16, 6)) / 4.0     # many exact ties
        rec["m_robot"] = r
        rec["m_pose"] = rng.integers(0, 50, (16, 6))
        rec["mono_inliers"] = rng.integers(0, 99, (16, 6))
        for b_ in range(16):                                          # per-rank lists arrive sorted, as the library emits them
            o = np.lexsort((rec["m_pose"][b_, :cntr[b_]], -rec["norm_bow_score"][b_, :cntr[b_]]))
            rec[b_, :cntr[b_]] = rec[b_, :cntr[b_]][o]
        blocks.append((rec, cntr))
    o1, c1 = shard.merge_records(blocks, 10)
    o2, c2 = shard.merge_records_native(blocks, 10)
    assert np.array_equal(c1, c2)
    for b_ in range(16):
        k1 = [(float(x["norm_bow_score"]), int(x["m_robot"]), int(x["m_pose"])) for x in o1[b_, :c1[b_]]]
        k2 = [(float(x["norm_bow_score"]), int(x["m_robot"]), int(x["m_pose"])) for x in o2[b_, :c2[b_]]]
        assert k1 == k2
    # a rank whose list is not in rank order (never produced by the library) takes the full-sort path
    rec0, cnt0 = blocks[0]
    for b_ in range(16):
        rec0[b_, :cnt0[b_]] = rec0[b_, :cnt0[b_]][::-1]
    o1, c1 = shard.merge_records(blocks, 10)
    o2, c2 = shard.merge_records_native(blocks, 10)
    assert np.array_equal(c1, c2)
    for b_ in range(16):
        k1 = [(float(x["norm_bow_score"]), int(x["m_robot"]), int(x["m_pose"])) for x in o1[b_, :c1[b_]]]
        k2 = [(float(x["norm_bow_score"]), int(x["m_robot"]), int(x["m_pose"])) for x in o2[b_, :c2[b_]]]
        assert k1 == k2
