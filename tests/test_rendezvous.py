"""kml/rendezvous.py — the TCP rendezvous bench.py and the sharded tests use instead of
torch.distributed: gather / broadcast / barrier / max over three local processes."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, key, q):
    sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
    from kml.rendezvous import Rendezvous
    r = Rendezvous(rank=rank, world=world, addr="127.0.0.1", port=port, key=key)
    g = r.gather({"rank": rank, "blob": bytes([rank]) * 100000})
    ok = [x["rank"] for x in g] == list(range(world)) and all(len(x["blob"]) == 100000 for x in g)
    uid = r.broadcast(b"u" * 128 if rank == 0 else None)
    ok = ok and uid == b"u" * 128
    r.barrier()
    ok = ok and r.max(float(rank)) == float(world - 1)
    ok = ok and r.broadcast("from2" if rank == 2 else None, src=2) == "from2"
    r.close()
    q.put((rank, ok))


def test_three_process_rendezvous():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    key = "test_%d" % os.getpid()
    procs = [ctx.Process(target=_worker, args=(r, 3, 29123, key, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True), (2, True)]


def test_single_process_rendezvous_is_a_no_op():
    sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))
    from kml.rendezvous import Rendezvous
    r = Rendezvous(rank=0, world=1)
    assert r.gather(5) == [5] and r.broadcast("x") == "x" and r.max(2.5) == 2.5
    r.barrier()
    r.close()
