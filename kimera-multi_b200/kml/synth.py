"""Seeded synthetic workload for the loop-closure hot path.

Test / benchmark input only — no part of the product calls this.  Everything
is derived from (master seed, tag, index) through numpy SeedSequence, so any
process regenerates identical bytes.

World model
  * P latent places.  Place p owns F ORB-256 prototype descriptors, F 3-D
    points (place frame, frustum [-5,5]^2 x [2,12] m) and F vocabulary words
    out of W = 10^6 (k=10, L=6 leaves).  With `alias=True` places p and
    p + P/2 share their WORDS (perceptual aliasing: the BoW scorer proposes
    them, geometric verification must reject them) but nothing else.
  * A keyframe observing place p keeps each prototype feature with prob 0.7
    (descriptor bits flipped with prob 1/16, word replaced by a random word
    with prob 0.1) and replaces the rest by random outliers; feature order is
    shuffled.  Camera pose: rotation <= 20 deg, translation <= 1 m.  Bearings
    carry 5e-4 rad noise, 3-D keypoints carry depth noise 0.005 z^2 and 10 %
    are zero (invalid depth).  15 % of the kept features are "moved": their
    descriptor still matches but bearing / 3-D point are random, which puts
    the RANSAC inlier ratio of a true pair near 0.7 (tens of iterations).
  * BoW vector = TF-IDF with a seeded log-uniform IDF in [0.5, 8],
    L1-normalised, rounded to float32 (the ROS wire type).
"""
import numpy as np

MASTER_SEED = 20241014
W_WORDS = 10 ** 6


def _rng(*key):
    return np.random.default_rng([MASTER_SEED] + [int(k) for k in key])


class World:
    def __init__(self, n_places, F=500, alias=True, seed_tag=0):
        self.P, self.F, self.alias, self.tag = int(n_places), int(F), bool(alias), int(seed_tag)
        r = _rng(1, seed_tag)
        P, F = self.P, self.F
        self.idf = np.exp(r.uniform(np.log(0.5), np.log(8.0), W_WORDS))
        self.proto_desc = r.integers(0, 256, (P, F, 32), dtype=np.uint8)
        words = r.integers(0, W_WORDS, (P, F), dtype=np.int64)
        if alias and P >= 2:
            half = P // 2
            words[half:2 * half] = words[:half]
        self.proto_words = words
        pts = np.empty((P, F, 3))
        pts[..., 0] = r.uniform(-5, 5, (P, F))
        pts[..., 1] = r.uniform(-5, 5, (P, F))
        pts[..., 2] = r.uniform(2, 12, (P, F))
        self.proto_pts = pts

    def place_of(self, robot, pose):
        return (int(pose) + 17 * int(robot)) % self.P

    # ------------------------------------------------------------------
    def frames(self, robot, poses, places=None, key=0):
        """Generate keyframes.  Returns dict with desc [n,F,32] u8, bearings
        [n,F,3], points [n,F,3], bow (off, ids, vals), R [n,3,3], t [n,3]
        (camera pose in the place frame: X_place = R X_cam + t)."""
        poses = np.asarray(poses, dtype=np.int64)
        n, F = len(poses), self.F
        if places is None:
            places = np.array([self.place_of(robot, p) for p in poses], dtype=np.int64)
        places = np.asarray(places, dtype=np.int64)
        r = _rng(2, self.tag, robot, key, int(poses[0]) if n else 0, n)
        keep = r.random((n, F)) < 0.7
        # descriptors: prototype with ~1/16 of the bits flipped, or random
        flip = (r.integers(0, 256, (n, F, 32), dtype=np.uint8) & r.integers(0, 256, (n, F, 32), dtype=np.uint8)
                & r.integers(0, 256, (n, F, 32), dtype=np.uint8) & r.integers(0, 256, (n, F, 32), dtype=np.uint8))
        desc = self.proto_desc[places] ^ flip
        rnd_desc = r.integers(0, 256, (n, F, 32), dtype=np.uint8)
        desc = np.where(keep[..., None], desc, rnd_desc)
        # words
        words = self.proto_words[places].copy()
        rw = r.integers(0, W_WORDS, (n, F), dtype=np.int64)
        swap = (~keep) | (r.random((n, F)) < 0.1)
        words = np.where(swap, rw, words)
        # camera pose
        axis = r.normal(size=(n, 3))
        axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        ang = r.uniform(0, np.deg2rad(20.0), n)
        R = _rodrigues(axis * ang[:, None])
        t = r.uniform(-1, 1, (n, 3)) / np.sqrt(3.0)
        X = self.proto_pts[places]                               # [n,F,3] place frame
        Xc = np.einsum("nji,nfj->nfi", R, X - t[:, None, :])     # R^T (X - t)
        depth = np.linalg.norm(Xc, axis=2, keepdims=True)
        dirs = Xc / depth
        noise = r.normal(size=(n, F, 3)) * 5e-4
        bear = dirs + noise
        bear /= np.linalg.norm(bear, axis=2, keepdims=True)
        dn = depth + r.normal(size=(n, F, 1)) * 0.005 * Xc[..., 2:3] ** 2
        pts = bear * dn
        # outliers: random directions in the forward cone, random depth
        ob = r.normal(size=(n, F, 3)) * 0.4
        ob[..., 2] = 1.0
        ob /= np.linalg.norm(ob, axis=2, keepdims=True)
        op = ob * r.uniform(2, 12, (n, F, 1))
        # 15 % of the kept features sit on "moved" structure: the descriptor
        # still matches but the geometry is an outlier for RANSAC
        moved = r.random((n, F)) < 0.15
        geo_ok = keep & ~moved
        bear = np.where(geo_ok[..., None], bear, ob)
        pts = np.where(geo_ok[..., None], pts, op)
        pts = np.where((r.random((n, F)) < 0.1)[..., None], 0.0, pts)
        # shuffle feature order per frame
        perm = np.argsort(r.random((n, F)), axis=1)
        desc = np.take_along_axis(desc, perm[..., None], axis=1)
        bear = np.take_along_axis(bear, perm[..., None], axis=1)
        pts = np.take_along_axis(pts, perm[..., None], axis=1)
        off, ids, vals = self._bow(words)
        return dict(robot=int(robot), poses=poses.astype(np.uint64), places=places,
                    desc=np.ascontiguousarray(desc), bearings=np.ascontiguousarray(bear),
                    points=np.ascontiguousarray(pts), bow_off=off, bow_ids=ids, bow_vals=vals,
                    R=R, t=t)

    def _bow(self, words):
        n = words.shape[0]
        off = np.zeros(n + 1, np.int64)
        ids_l, vals_l = [], []
        for i in range(n):
            w, cnt = np.unique(words[i], return_counts=True)
            v = cnt * self.idf[w]
            v = (v / v.sum()).astype(np.float32)
            ids_l.append(w.astype(np.uint32))
            vals_l.append(v)
            off[i + 1] = off[i] + len(w)
        ids = np.concatenate(ids_l) if ids_l else np.zeros(0, np.uint32)
        vals = np.concatenate(vals_l) if vals_l else np.zeros(0, np.float32)
        return off, ids, vals


def _rodrigues(rv):
    th = np.linalg.norm(rv, axis=1)
    k = rv / np.maximum(th, 1e-300)[:, None]
    K = np.zeros((len(rv), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def relative_pose(fq, iq, fm, im):
    """Ground truth x_q = R x_m + t between frame iq of fq and frame im of fm
    (only meaningful when both observe the same place)."""
    Rq, tq, Rm, tm = fq["R"][iq], fq["t"][iq], fm["R"][im], fm["t"][im]
    return Rq.T @ Rm, Rq.T @ (tm - tq)


def build_database(world, robots, n_keyframes, chunk=512):
    """Yield per-robot chunks of database keyframes (poses 0..n_keyframes-1)."""
    for r in robots:
        for s in range(0, n_keyframes, chunk):
            poses = np.arange(s, min(s + chunk, n_keyframes))
            yield world.frames(r, poses)


def make_queries(world, B, n_keyframes, n_robots, key=0, robots=None):
    """B held-out revisit queries: query b is a new keyframe (pose
    n_keyframes + 2b + 1 of a round-robin robot) of a random place; its
    "previous BoW" is another fresh observation of the same place."""
    r = _rng(3, world.tag, key)
    places = r.integers(0, world.P, B)
    q_robot = (np.arange(B) % n_robots).astype(np.uint64) if robots is None else np.asarray(robots, np.uint64)
    q_pose = (n_keyframes + 2 * np.arange(B) + 1).astype(np.uint64)
    fq = world.frames(1000 + key, q_pose, places=places, key=1)
    fp = world.frames(1000 + key, q_pose - 1, places=places, key=2)
    return dict(q_robot=q_robot, q_pose=q_pose, places=places, frames=fq, prev=fp)


def make_vocabulary(k=10, L=6, key=0):
    """Synthetic DBoW2 vocabulary tree: random 256-bit node descriptors in
    breadth-first order (level 1 first) and a log-uniform IDF weight in [0.5, 8] per leaf;
    2 % of the words get weight 0 (DBoW2 drops words seen in every training image)."""
    r = _rng(4, key)
    n_nodes = sum(k ** l for l in range(1, L + 1))
    nodes = r.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
    w = np.exp(r.uniform(np.log(0.5), np.log(8.0), k ** L))
    w[r.random(k ** L) < 0.02] = 0.0
    return nodes, w
