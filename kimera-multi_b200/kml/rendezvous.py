"""A few-hundred-byte rendezvous for the one-process-per-GPU launch (torchrun or any launcher that
sets RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT): rank 0 listens on a TCP socket, every rank
sends one pickled object per collective and receives the list of all of them.  That is all the
sharded query needs around NCCL — the ncclUniqueId broadcast, barriers around the timed region and
the max-over-ranks of a time — so neither the library nor bench.py imports torch.distributed.

Rank 0 binds an ephemeral port and publishes it in a file keyed by (MASTER_PORT, launcher pid):
MASTER_PORT itself belongs to the launcher's own store.
"""
import os
import pickle
import socket
import struct
import time


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        hdr += chunk
    n = struct.unpack("<Q", hdr)[0]
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return pickle.loads(bytes(buf))


class Rendezvous:
    """gather / broadcast / barrier / max over `world` processes of one box."""

    def __init__(self, rank=None, world=None, addr=None, port=None, key=None, timeout=600.0):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        self.addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
        self.timeout = timeout
        self.peers = []
        self.sock = None
        if self.world == 1:
            return
        base = int(os.environ.get("MASTER_PORT", "29500")) if port is None else int(port)
        key = key or os.environ.get("TORCHELASTIC_RUN_ID", "") + "_" + str(os.getppid())
        path = "/tmp/kml_rdzv_%d_%s" % (base, "".join(c if c.isalnum() else "_" for c in key))
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((self.addr if self.addr != "localhost" else "127.0.0.1", 0))
            srv.listen(self.world)
            srv.settimeout(timeout)
            with open(path + ".tmp", "w") as f:
                f.write(str(srv.getsockname()[1]))
            os.replace(path + ".tmp", path)
            got = {}
            while len(got) < self.world - 1:
                c, _ = srv.accept()
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                c.settimeout(timeout)
                got[_recv(c)] = c
            self.peers = [got[r] for r in range(1, self.world)]
            srv.close()
            try:
                os.unlink(path)
            except OSError:
                pass
        else:
            t0 = time.time()
            while True:
                try:
                    with open(path) as f:
                        p = int(f.read().strip())
                    s = socket.create_connection((self.addr, p), timeout=timeout)
                    break
                except (OSError, ValueError):
                    if time.time() - t0 > timeout:
                        raise TimeoutError("rendezvous: rank 0 never published %s" % path)
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            _send(s, self.rank)
            self.sock = s

    def gather(self, obj):
        """every rank contributes `obj`; every rank receives the list indexed by rank"""
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            objs = [obj] + [_recv(c) for c in self.peers]
            for c in self.peers:
                _send(c, objs)
            return objs
        _send(self.sock, obj)
        return _recv(self.sock)

    def broadcast(self, obj, src=0):
        return self.gather(obj if self.rank == src else None)[src]

    def barrier(self):
        self.gather(None)

    def max(self, x):
        return max(self.gather(x))

    def close(self):
        for c in self.peers:
            c.close()
        if self.sock:
            self.sock.close()
        self.peers, self.sock = [], None
