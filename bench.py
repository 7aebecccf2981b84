#!/usr/bin/env python
"""bench.py — loop-closure queries/s on B200 (BASELINE.json metric).

Workload (config C2): 6 robot databases x 5 000 keyframes per
GPU (500 ORB-256 features / keyframe, synthetic k=10 L=6 vocabulary), queries
in batches of 256, top_k_verify = 16 candidates verified per query
(kNN + Lowe 0.9 -> mono 5-pt RANSAC -> stereo Arun RANSAC).  One "step" = one
batch of 256 queries through the whole hot path.

  value : queries/s with the batch already resident in HBM (kml_query_batch_run),
          timed on the device with CUDA events on the library's stream.
  e2e   : the same through kml_query_batch with HOST buffers (H2D of the
          batch + D2H of the records inside the timed region).
  N > 1 : one process per GPU, databases sharded by robot (each rank holds its
          own 6 x 5 000 shard), the batch is replicated on every rank and the
          per-shard records are merged by ONE ncclAllGather per step; value =
          (N * 256 shard-queries) / max-over-ranks step time  ("weak").

`--impl reference` times the CPU oracle (scalar C++ restatement of the
reference algorithms, all host threads) on bounded samples of the same
workload.
"""
import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kimera-multi_b200"))

import numpy as np  # noqa: E402

N_ROBOTS = 6          # per GPU; --workload c5 sets 1 x 50 000 (BASELINE.json configs[4])
N_KEYFRAMES = 5000
BATCH = 256
WORKLOAD = "C2"
F = 500
METRIC = "loop_closure_queries_per_s"
UNIT = "queries/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_world(rank, log):
    from kml import synth
    t0 = time.time()
    world = synth.World(N_KEYFRAMES // 4, F=F)
    robots = [rank * N_ROBOTS + r for r in range(N_ROBOTS)]
    log("world ready in %.1fs; shard robots %s" % (time.time() - t0, robots))
    return world, robots


def fill_detector(det, world, robots, log):
    from kml import synth
    t0 = time.time()
    for ch in synth.build_database(world, robots, N_KEYFRAMES, chunk=1000):
        det.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
        det.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
    log("database of %d x %d keyframes resident in %.1fs" % (len(robots), N_KEYFRAMES, time.time() - t0))


def make_batches(world, n, n_robots_total, B=BATCH):
    from kml import synth
    out = []
    for k in range(n):
        q = synth.make_queries(world, B, N_KEYFRAMES, n_robots_total, key=k)
        fq, fp = q["frames"], q["prev"]
        out.append((q["q_robot"], q["q_pose"], fq["bow_off"], fq["bow_ids"], fq["bow_vals"],
                    fp["bow_off"], fp["bow_ids"], fp["bow_vals"], fq["desc"], fq["bearings"],
                    fq["points"]))
    return out


def batch_bytes(b):
    return int(sum(np.asarray(x).nbytes for x in b))


def run_reference(args, rank, world_size, log):
    """CPU arm: the oracle (scalar C++ port of DBoW2 + BFMatcher + OpenGV paths)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kml_oracle as ko
    world, robots = build_world(0, log)
    lcd = ko.LoopClosureDetector()
    from kml import synth
    t0 = time.time()
    for ch in synth.build_database(world, robots, N_KEYFRAMES, chunk=1000):
        for i, p in enumerate(ch["poses"]):
            o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
            lcd.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
            lcd.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
    log("oracle database ready in %.1fs" % (time.time() - t0))
    threads = len(os.sched_getaffinity(0))  # torchrun pins OMP_NUM_THREADS=1: ask for every core explicitly
    sample = BATCH if threads >= 4 else 64  # a step = the whole 256-query batch of the GPU arm
    batches = make_batches(world, args.steps + args.warmup, N_ROBOTS, B=sample)
    times = []
    for i, b in enumerate(batches):
        t0 = time.perf_counter()
        lcd.query_batch(*b, threads=threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
        "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d-query batches of the %s workload per step, %d OpenMP threads" % (sample, WORKLOAD, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "%s: %d robot DBs x %d keyframes per GPU, %d ORB-256 features/keyframe, "
                        "%d-query batches, top_k_verify 16, lowe 0.9, mono 5-pt + stereo Arun RANSAC "
                        "(max 1000 it, p 0.995)" % (WORKLOAD, N_ROBOTS, N_KEYFRAMES, F, BATCH),
            "global_batch": BATCH, "n_robots_per_gpu": N_ROBOTS, "keyframes_per_robot": N_KEYFRAMES,
            "parallelism": "robot-sharded x%d, replicated query batch + 1 ncclAllGather/step" % n if n > 1 else "single GPU",
            "l2": "256 MiB device memset before the timed region; each step re-reads ~200 MB (touched postings, 4 096 candidate "
                  "frames, the batch) per lane, the query lanes interleaved, against a 126 MB L2; the e2e arm uses a distinct batch every step"}


def main():
    global N_ROBOTS, N_KEYFRAMES, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kml", choices=["kml", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0,
                    help="query batches kept in flight per GPU (0 = default, 4)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: 6 robot DBs x 5 000 keyframes per GPU (the headline config); "
                         "c5: one robot DB of 50 000 keyframes per GPU (BASELINE.json configs[4])")
    ap.add_argument("--watchdog-s", type=int, default=1500,
                    help="dump every thread's stack and exit if the run has not finished by then (0 = off)")
    args = ap.parse_args()
    if args.watchdog_s > 0:
        # a wedged collective or lane thread must end the process, not hold the GPU box
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    if args.workload == "c5":
        N_ROBOTS, N_KEYFRAMES, WORKLOAD = 1, 50000, "C5"
    args.warmup = max(args.warmup, 3) if args.impl == "kml" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if world_size > 1 and args.lanes > 4:
        # every lane launches its all-gather on its own stream in its own order; with more streams
        # than hardware queues (8 by default) two lanes can share a queue, and a rank whose lane-0
        # collective sits behind its lane-5 collective then waits on a peer that queued them the
        # other way round (8 GPUs x 6 lanes did not finish in 200 s; <= 4 lanes never stalled).
        # The driver reads this when the CUDA context is created.
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

    def log(msg):
        print("[bench r%d] %s" % (rank, msg), file=sys.stderr, flush=True)

    if args.impl == "reference":
        run_reference(args, rank, world_size, log)
        return

    import kml
    dist = None
    if world_size > 1:
        # NCCL prints its version / debug lines on stdout by default: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world_size)

    def barrier():
        if dist is not None:
            dist.barrier()

    kml.build()
    det = kml.LoopClosureDetector(device=local_rank)
    n_lanes = args.lanes if args.lanes > 0 else 4
    lanes = [det] + [det.create_lane() for _ in range(n_lanes - 1)]
    if world_size > 1:
        # one communicator per lane: lane i of every rank forms its own all-gather group, so the
        # lanes' collectives never have to be ordered against each other
        uids = [[kml.LoopClosureDetector.comm_unique_id() for _ in lanes] if rank == 0 else None]
        dist.broadcast_object_list(uids, src=0)
        # NCCL prints its version line on stdout while a communicator is created: point fd 1 at
        # stderr for the duration so that stdout carries nothing but the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            for ln, uid in zip(lanes, uids[0]):
                ln.comm_init(world_size, rank, uid)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    world, robots = build_world(rank, log)
    fill_detector(det, world, robots, log)
    n_batches = args.steps + args.warmup
    batches = make_batches(world, min(n_batches, 6), N_ROBOTS * world_size)
    sharded = world_size > 1
    popc_peak = det.peak_popc()
    fp64_peak = det.peak_fp64()

    # ---------------- resident arm: device-timed steps
    # Two query lanes (kml_create_lane) keep two batches in flight on the GPU: the tail rounds
    # of one batch's RANSAC and the host-side candidate selection overlap the other batch.
    for li, ln in enumerate(lanes):
        for i in range(args.warmup):
            ln.query_batch_upload(*batches[(li + i) % len(batches)])
            ln.query_batch_run(sharded=sharded)
        ln.query_batch_upload(*batches[li % len(batches)])     # this lane's resident batch
    barrier()
    # clocks / throttle reasons are sampled on rank 0's GPU only: eight nvidia-smi pollers on one
    # box compete with the lane threads for the host cores
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = sum(ln.stats().kernel_launches for ln in lanes)
    last = {}
    lock = threading.Lock()

    def resident_worker(ln, n):
        for _ in range(n):
            out_, counts_ = ln.query_batch_run(sharded=sharded)
            with lock:
                last["out"], last["counts"] = out_, counts_

    shares = [args.steps // n_lanes + (1 if i < args.steps % n_lanes else 0) for i in range(n_lanes)]
    det.flush_l2()
    t0 = time.perf_counter()
    det.timer_begin()
    threads = [threading.Thread(target=resident_worker, args=(ln, n)) for ln, n in zip(lanes, shares)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    step_ms = det.timer_end(lanes)            # CUDA events: begin on lane 0, end after every lane's last kernel
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = sum(ln.stats().kernel_launches for ln in lanes) - l0
    clocks = sampler.stop() if sampler else None
    barrier()
    out, counts = last["out"], last["counts"]
    if dist is not None:
        import torch
        t = torch.tensor([step_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t[0])
    value = world_size * BATCH * args.steps / (step_ms * 1e-3)

    # ---------------- e2e arm: host buffers in, records out, every step (distinct batch per step)
    if not sharded:
        for ln in lanes:
            ln.query_batch(*batches[0])
    barrier()

    def e2e_worker(ln, idxs):
        for i in idxs:
            b = batches[(args.warmup + i) % len(batches)]
            if sharded:
                ln.query_batch_upload(*b)
                ln.query_batch_run(sharded=True)
            else:
                ln.query_batch(*b)

    det.flush_l2()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=e2e_worker, args=(ln, range(li, args.steps, n_lanes)))
               for li, ln in enumerate(lanes)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    e2e_t = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_t], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t[0])
    e2e_value = world_size * BATCH * args.steps / e2e_t
    h2d = batch_bytes(batches[0])
    d2h = int(out.nbytes + counts.nbytes)

    # ---------------- attribution pass: per-stage device time of ONE lane running alone
    # (with two lanes in flight the per-stage CUDA events of a lane also span the other lane's
    # kernels, so stage shares and rooflines come from this single-lane pass, outside the timed region)
    K = min(args.steps, 10)
    stage = {"bow": 0.0, "match": 0.0, "mono": 0.0, "stereo": 0.0}
    acc = {"postings": 0, "hyp_m": 0, "hyp_s": 0, "pairs": 0, "res_m": 0, "res_s": 0}
    det.query_batch_upload(*batches[0])
    barrier()
    for _ in range(K):
        det.query_batch_run(sharded=sharded)
        st = det.stats()
        stage["bow"] += st.ms_bow; stage["match"] += st.ms_match
        stage["mono"] += st.ms_mono; stage["stereo"] += st.ms_stereo
        acc["postings"] += st.bow_postings_last; acc["hyp_m"] += st.mono_hypotheses_last
        acc["hyp_s"] += st.stereo_hypotheses_last; acc["pairs"] += st.pairs_last
        acc["res_m"] += st.mono_residuals_last; acc["res_s"] += st.stereo_residuals_last
    postings, hyp_m, hyp_s, pairs = acc["postings"], acc["hyp_m"], acc["hyp_s"], acc["pairs"]
    res_m, res_s = acc["res_m"], acc["res_s"]
    barrier()

    # ---------------- rooflines (algorithmic work / device time per stage)
    hbm_peak, peak_src = load_peaks()
    compares = pairs * F * F
    roof = {
        "bow_scan": {"bound": "hbm", "achieved": postings * 8 / (stage["bow"] * 1e-3) / 1e9 if stage["bow"] else None,
                     "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                     "algorithmic_bytes_per_step": postings * 8 / K, "ms_per_step": stage["bow"] / K},
        # algorithmic: 8 POPC32 per 256-bit compare; the kernel executes 5 (three carry-save adders fold
        # seven XOR words into two weight-1 and three weight-2 words first), so `achieved` can exceed
        # the pipe's peak; `frac` is the utilisation of the POPC pipe by what is actually executed
        "hamming_knn": {"bound": "popc", "achieved": compares * 8 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                        "executed": compares * 5 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                        "peak": popc_peak / 1e12, "unit": "T POPC32/s", "peak_source": "measured (kml_peak_popc, same run)",
                        "compares_per_step": compares / K, "ms_per_step": stage["match"] / K},
        # algorithmic flops of the reference loop: per consumed hypothesis the minimal
        # solver (mono 33 kflop, stereo 1.5 kflop) plus one residual per correspondence (mono 95, stereo 27 flop)
        "mono_ransac": {"bound": "fp64", "hypotheses_per_step": hyp_m / K, "residuals_per_step": res_m / K,
                        "ms_per_step": stage["mono"] / K,
                        "achieved": (hyp_m * 33e3 + res_m * 95.0) / (stage["mono"] * 1e-3) / 1e12 if stage["mono"] else None,
                        "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "peak_source": "measured (kml_peak_fp64, same run)"},
        "stereo_ransac": {"bound": "fp64", "hypotheses_per_step": hyp_s / K, "residuals_per_step": res_s / K,
                          "ms_per_step": stage["stereo"] / K,
                          "achieved": (hyp_s * 1.5e3 + res_s * 27.0) / (stage["stereo"] * 1e-3) / 1e12 if stage["stereo"] else None,
                          "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "peak_source": "measured (kml_peak_fp64, same run)"},
    }
    for k in ("bow_scan", "hamming_knn", "mono_ransac", "stereo_ransac"):
        if roof[k]["achieved"]:
            roof[k]["frac"] = roof[k]["achieved"] / roof[k]["peak"]
    if roof["hamming_knn"].get("executed"):
        roof["hamming_knn"]["frac_algorithmic"] = roof["hamming_knn"]["frac"]
        roof["hamming_knn"]["frac"] = roof["hamming_knn"]["executed"] / roof["hamming_knn"]["peak"]
    # flop model: 5-pt hypothesis ~ 60 kflop solver + 8x4x10 scoring; residual 95 flop
    dominant = max(("bow", "match", "mono", "stereo"), key=lambda s: stage[s])
    dom_key = {"bow": "bow_scan", "match": "hamming_knn", "mono": "mono_ransac", "stereo": "stereo_ransac"}[dominant]
    roofline = dict(roof[dom_key])
    roofline["kernel"] = dom_key
    roofline.setdefault("achieved", None)
    roofline.setdefault("frac", None)
    roofline["traffic"] = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64-fixed/f64", "data": "synthetic",
        "config": workload_config(world_size), "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roofline, "rooflines": roof,
        "stage_ms_per_step": {k: v / K for k, v in stage.items()},
        "stage_ms_source": "single-lane attribution pass of %d steps after the timed region" % K,
        "wall_ms_per_step": wall_ms / args.steps, "lanes": n_lanes,
        "loop_closures_last_step": int((out["status"][counts[:, None] > np.arange(out.shape[1])[None]] == 0).sum()),
    }
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(world, robots, log)
    if rank == 0:
        print(json.dumps(line), flush=True)
    det.close()
    if dist is not None:
        dist.destroy_process_group()


def _oracle_rate(ko, world, robots, threads, sample, log, tag):
    """queries/s of one `sample`-query batch on a freshly filled oracle detector of module `ko`"""
    from kml import synth
    lcd = ko.LoopClosureDetector()
    t0 = time.time()
    for ch in synth.build_database(world, robots, N_KEYFRAMES, chunk=1000):
        for i, p in enumerate(ch["poses"]):
            o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
            lcd.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
            lcd.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])
    log("oracle database (%s) ready in %.1fs" % (tag, time.time() - t0))
    b = make_batches(world, 1, N_ROBOTS, B=sample)[0]
    lcd.query_batch(*b, threads=threads)  # warm-up
    t0 = time.perf_counter()
    lcd.query_batch(*b, threads=threads)
    dt = time.perf_counter() - t0
    del lcd
    return sample / dt, dt


NATIVE_FLAGS = "-O3 -march=native -std=c++17 -fPIC -ffp-contract=off -fopenmp"


def _native_oracle(log):
    """The oracle rebuilt on THIS host with -O3 -march=native (the "fair best CPU" beside
    the reference's own -O2 / no -march build), into a temporary directory, loaded as a second copy
    of the binding module.  None if the compiler is not there."""
    import importlib.util
    import tempfile
    odir = os.path.join(ROOT, "oracle")
    try:
        so = os.path.join(tempfile.mkdtemp(prefix="kml_oracle_native_"), "libkml_oracle_native.so")
        subprocess.run(["g++"] + NATIVE_FLAGS.split() + ["-shared", "-o", so, os.path.join(odir, "src", "oracle.cpp")],
                       check=True, capture_output=True, timeout=600)
        spec = importlib.util.spec_from_file_location("kml_oracle_native", os.path.join(odir, "kml_oracle.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod._LIB_PATH = so
        mod.lib()
        return mod
    except Exception as e:  # noqa: BLE001 - a baseline extra, never fatal
        log("native oracle build skipped: %r" % (e,))
        return None


def cpu_baseline(world, robots, log):
    """The oracle timed on this box's host cores on a bounded sample (rank 0, N=1): built like the
    reference (-O2, no -march=native), and once more with -O3 -march=native as the fair best CPU."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kml_oracle as ko
    threads = len(os.sched_getaffinity(0))
    sample = BATCH if threads >= 4 else 64  # the whole 256-query batch: ~15-20 core-seconds
    value, dt = _oracle_rate(ko, world, robots, threads, sample, log, "-O2")
    out = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d queries (one batch) of the %s workload, %d OpenMP threads, %.2fs" % (sample, WORKLOAD, threads, dt)}
    native = _native_oracle(log)
    if native is not None:
        v2, dt2 = _oracle_rate(native, world, robots, threads, sample, log, "-O3 -march=native")
        out["fair_best"] = {"value": v2, "unit": UNIT, "cores": threads, "flags": NATIVE_FLAGS,
                            "sample": "same batch, %.2fs" % dt2}
    return out


if __name__ == "__main__":
    main()
