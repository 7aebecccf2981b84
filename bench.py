#!/usr/bin/env python
"""bench.py — loop-closure queries/s on B200 (BASELINE.json metric).

Workloads:
  N = 1  config C2 (BASELINE.json configs[1], the one the metric is quoted on): 6 robot databases x
         5 000 keyframes on one GPU, 500 ORB-256 features / keyframe, synthetic k=10 L=6
         vocabulary, queries in batches of 256, top_k_verify = 16 candidates verified per query
         (kNN + Lowe 0.9 -> mono 5-pt RANSAC -> stereo Arun RANSAC).
  N > 1  config C5 (configs[4]): one robot database of 50 000 keyframes per GPU, the query batch
         replicated on every rank, per-shard records merged by ONE ncclAllGather per step + a
         device merge.  `--workload` overrides either default.
One "step" = one batch of 256 queries through the whole hot path.

  value : with the batch already resident in HBM (kml_query_batch_run / _sharded), timed on the
          device with CUDA events; K steps per timed block, blocks repeated until the timed
          region is >= 1 s, median block reported (max over ranks per block).
          N > 1: value = N * 256 * K / t  counts (query, shard) units — every GPU processes every
          query against its own shard (weak scaling in database size); `distinct_queries_per_s`
          is value / N.
  e2e   : the same through kml_query_batch (N = 1) or upload + sharded run (N > 1) with HOST
          buffers in pinned memory: H2D of the batch + D2H of the records inside the timed region.

`--impl reference` times the CPU oracle (scalar C++ restatement of the reference algorithms, built
with the reference's flags, all host threads) on the same workload.  No torch anywhere: ranks
meet through kml/rendezvous.py (a TCP socket) and NCCL is driven by the library itself.
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

N_ROBOTS = 6
N_KEYFRAMES = 5000
BATCH = 256
WORKLOAD = "C2"
F = 500
METRIC = "loop_closure_queries_per_s"
UNIT = "queries/s"
MIN_TIMED_S = 1.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", float(d.get("bf16_tflops", 1660.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1660.0


def load_traffic():
    """dram__bytes per launch of the BoW and Hamming kernels from the committed ncu capture"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


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


def fill_oracle(lcd, world, robots, n_keyframes):
    from kml import synth
    for ch in synth.build_database(world, robots, n_keyframes, chunk=1000):
        for i, p in enumerate(ch["poses"]):
            o0, o1 = ch["bow_off"][i], ch["bow_off"][i + 1]
            lcd.addBowVector(ch["robot"], int(p), ch["bow_ids"][o0:o1], ch["bow_vals"][o0:o1])
            lcd.addVLCFrame(ch["robot"], int(p), ch["desc"][i], ch["bearings"][i], ch["points"][i])


def make_batches(world, n, n_robots_total, B=BATCH, pin=None):
    from kml import synth
    out = []
    for k in range(n):
        q = synth.make_queries(world, B, N_KEYFRAMES, n_robots_total, key=k)
        fq, fp = q["frames"], q["prev"]
        b = (q["q_robot"], q["q_pose"], fq["bow_off"], fq["bow_ids"], fq["bow_vals"],
             fp["bow_off"], fp["bow_ids"], fp["bow_vals"], fq["desc"], fq["bearings"], fq["points"])
        if pin is not None:   # host buffers of the e2e arm: page-locked (kml_host_alloc)
            b = tuple(pin(np.ascontiguousarray(x)) for x in b)
        out.append(b)
    return out


def batch_bytes(b):
    return int(sum(np.asarray(x).nbytes for x in b))


def workload_config(n, sharded):
    return {"workload": "%s: %d robot DB%s x %d keyframes per GPU, %d ORB-256 features/keyframe, "
                        "%d-query batches, top_k_verify 16, lowe 0.9, mono 5-pt + stereo Arun RANSAC "
                        "(max 1000 it, p 0.995)" % (WORKLOAD, N_ROBOTS, "" if N_ROBOTS == 1 else "s", N_KEYFRAMES, F, BATCH),
            "global_batch": BATCH, "n_robots_per_gpu": N_ROBOTS, "keyframes_per_robot": N_KEYFRAMES,
            "parallelism": ("robot-sharded x%d, replicated query batch + 1 in-place ncclAllGather/step + device merge" % n)
                           if sharded else "single GPU",
            "l2": "256 MiB device memset (> 126 MB L2) before every timed block; each step re-reads ~200 MB (touched "
                  "postings, <= 4 096 candidate frames, the batch) per lane, the query lanes interleaved; the e2e arm "
                  "uses a distinct batch every step"}


def run_reference(args, rank, log):
    """CPU arm: the oracle (scalar C++ port of DBoW2 + BFMatcher + OpenGV paths), reference build flags."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kml_oracle as ko
    world, robots = build_world(0, log)
    lcd = ko.LoopClosureDetector()
    t0 = time.time()
    fill_oracle(lcd, world, robots, N_KEYFRAMES)
    log("oracle database ready in %.1fs" % (time.time() - t0))
    threads = len(os.sched_getaffinity(0))  # torchrun pins OMP_NUM_THREADS=1: ask for every core explicitly
    sample = BATCH if threads >= 4 else 64  # a step = the whole 256-query batch of the GPU arm
    batches = make_batches(world, args.steps + args.warmup, N_ROBOTS * max(args.gpus, 1), B=sample)
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
        "config": workload_config(1, False),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "flags": "-O2 -ffp-contract=off -fopenmp, no -march=native (the reference's own build, README.md:116)",
                         "sample": "%d-query batches of ONE %s shard per step, %d OpenMP threads" % (sample, WORKLOAD, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def records_equal(a, ca, b, cb):
    if not np.array_equal(ca, cb):
        return False
    for q in range(len(ca)):
        if a[q, :ca[q]].tobytes() != b[q, :cb[q]].tobytes():
            return False
    return True


def main():
    global N_ROBOTS, N_KEYFRAMES, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kml", choices=["kml", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 side measurements")
    ap.add_argument("--lanes", type=int, default=0,
                    help="query batches kept in flight per GPU (0 = default, 4)")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c5"],
                    help="auto: c2 on one GPU (6 robot DBs x 5 000 keyframes, the headline config), c5 on several "
                         "(one robot DB of 50 000 keyframes per GPU, BASELINE.json configs[4])")
    ap.add_argument("--min-timed-s", type=float, default=MIN_TIMED_S)
    ap.add_argument("--matcher-engine", type=int, default=int(os.environ.get("KML_MATCHER_ENGINE", "1")), choices=[0, 1],
                    help="kml_params.matcher_engine: 0 = POPC pipe, 1 = tensor cores (tcgen05 kind::i8)")
    ap.add_argument("--watchdog-s", type=int, default=1500,
                    help="dump every thread's stack and exit if the run has not finished by then (0 = off)")
    args = ap.parse_args()
    if args.watchdog_s > 0:
        # a wedged collective or lane thread must end the process, not hold the GPU box
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    wl = args.workload if args.workload != "auto" else ("c5" if max(world_size, args.gpus) > 1 else "c2")
    if wl == "c5":
        N_ROBOTS, N_KEYFRAMES, WORKLOAD = 1, 50000, "C5"
    args.warmup = max(args.warmup, 3) if args.impl == "kml" else max(args.warmup, 1)
    if world_size > 1 and args.lanes > 4:
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

    def log(msg):
        print("[bench r%d] %s" % (rank, msg), file=sys.stderr, flush=True)

    if args.impl == "reference":
        run_reference(args, rank, log)
        return

    import kml
    from kml import shard
    from kml.rendezvous import Rendezvous
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries nothing but the JSON line
    rdzv = Rendezvous()
    barrier = rdzv.barrier

    kml.build()
    prm = kml.default_params()
    prm.matcher_engine = args.matcher_engine
    det = kml.LoopClosureDetector(prm, device=local_rank)
    n_lanes = args.lanes if args.lanes > 0 else 4
    lanes = [det] + [det.create_lane() for _ in range(n_lanes - 1)]
    sharded = world_size > 1
    if sharded:
        # one communicator per lane: lane i of every rank forms its own all-gather group; the lanes'
        # collectives are submitted in the batches' global sequence order (kml_query_batch_sharded_seq)
        uids = rdzv.broadcast([kml.LoopClosureDetector.comm_unique_id() for _ in lanes] if rank == 0 else None)
        sys.stdout.flush()
        saved = os.dup(1)   # NCCL prints its version line on stdout while a communicator is created
        os.dup2(2, 1)
        try:
            for ln, uid in zip(lanes, uids):
                ln.comm_init(world_size, rank, uid)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    world, robots = build_world(rank, log)
    fill_detector(det, world, robots, log)
    batches = make_batches(world, 6, N_ROBOTS * world_size)
    popc_peak = det.peak_popc()
    fp64_peak = det.peak_fp64()
    seq = [0]           # next global batch sequence number (identical on every rank)

    def run_block(fn_per_lane):
        """one timed block: every lane thread runs its share of the K steps; returns after all joined"""
        threads = [threading.Thread(target=fn_per_lane, args=(li, ln)) for li, ln in enumerate(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()

    # ---------------- warm-up + sharded parity (N > 1): the NCCL-merged records of every rank must equal
    # the merge (kml/shard.py, the documented rule) of all ranks' local records
    sharded_parity = None
    for li, ln in enumerate(lanes):
        for i in range(args.warmup):
            ln.query_batch_upload(*batches[(li + i) % len(batches)])
            if sharded:
                det.comm_seq_reset(0)
                ln.query_batch_run(sharded=True, seq=0)
            else:
                ln.query_batch_run()
    if sharded:
        ok = True
        for k in range(2):
            det.query_batch_upload(*batches[k])
            loc, lcnt = det.query_batch_run(sharded=False)
            det.comm_seq_reset(0)
            mrg, mcnt = det.query_batch_run(sharded=True, seq=0)
            blocks = rdzv.gather((loc, lcnt))
            ref, rcnt = shard.merge_records(blocks, int(det.params.top_k_verify))
            ok = ok and records_equal(ref, rcnt, mrg, mcnt) and int(rcnt.sum()) > 0
        sharded_parity = all(rdzv.gather(bool(ok)))
        log("sharded parity (NCCL all-gather + device merge == merge of the gathered local records): %s" % sharded_parity)
    for li, ln in enumerate(lanes):
        ln.query_batch_upload(*batches[li % len(batches)])     # this lane's resident batch
    K = args.steps

    # ---------------- resident arm: device-timed blocks of K steps
    sampler = ClockSampler(local_rank) if rank == 0 else None   # rank 0's GPU only (host cores are scarce at N = 8)
    if sampler:
        sampler.start()
    last = {}
    lock = threading.Lock()

    def resident_lane(li, ln):
        for st in range(li, K, n_lanes):
            if sharded:
                out_, counts_ = ln.query_batch_run(sharded=True, seq=seq[0] + st)
            else:
                out_, counts_ = ln.query_batch_run()
            with lock:
                last["out"], last["counts"] = out_, counts_

    block_ms, launches_blk = [], 0
    total_s, blocks_done = 0.0, 0
    while True:
        barrier()
        det.flush_l2()
        if sharded:
            det.comm_seq_reset(0)
            seq[0] = 0
        l0 = sum(ln.stats().kernel_launches for ln in lanes)
        det.timer_begin()
        run_block(resident_lane)
        ms = det.timer_end(lanes)            # CUDA events: begin on lane 0, end after every lane's last kernel
        launches_blk = sum(ln.stats().kernel_launches for ln in lanes) - l0
        ms = rdzv.max(ms)                    # max over ranks
        block_ms.append(ms)
        total_s += ms * 1e-3
        blocks_done += 1
        if (total_s >= args.min_timed_s and blocks_done >= 3) or blocks_done >= 200:
            break
    clocks = sampler.stop() if sampler else None
    step_ms = float(np.median(block_ms))
    value = world_size * BATCH * K / (step_ms * 1e-3)
    out, counts = last["out"], last["counts"]

    # ---------------- the same shard on ONE GPU without the exchange (N > 1 only; rank 0 alone, others idle)
    single_same_workload = None
    if sharded:
        barrier()
        if rank == 0:
            def solo_lane(li, ln):
                for st in range(li, K, n_lanes):
                    ln.query_batch_run()
            ts = []
            for _ in range(3):
                det.flush_l2()
                det.timer_begin()
                run_block(solo_lane)
                ts.append(det.timer_end(lanes))
            single_same_workload = BATCH * K / (float(np.median(ts)) * 1e-3)
        barrier()

    # ---------------- e2e arm: pinned host buffers in, records out, every step (distinct batch per step)
    pbatches = make_batches(world, 6, N_ROBOTS * world_size, pin=kml.pinned_copy)

    def e2e_lane(li, ln):
        for st in range(li, K, n_lanes):
            b = pbatches[st % len(pbatches)]
            if sharded:
                ln.query_batch(*b, sharded=True, seq=seq[0] + st)
            else:
                ln.query_batch(*b)

    if sharded:
        det.comm_seq_reset(0)
    seq[0] = 0
    run_block(e2e_lane)      # warm-up of the e2e path
    e2e_ms, total_s, blocks_done = [], 0.0, 0
    while True:
        barrier()
        det.flush_l2()
        if sharded:
            det.comm_seq_reset(0)
        t0 = time.perf_counter()
        run_block(e2e_lane)
        dt = rdzv.max(time.perf_counter() - t0)
        e2e_ms.append(dt * 1e3)
        total_s += dt
        blocks_done += 1
        if (total_s >= args.min_timed_s and blocks_done >= 3) or blocks_done >= 200:
            break
    e2e_value = world_size * BATCH * K / (float(np.median(e2e_ms)) * 1e-3)
    h2d = batch_bytes(batches[0])
    d2h = int(out.nbytes + counts.nbytes)

    # ---------------- attribution pass: per-stage device time of ONE lane running alone
    KA = min(K, 10)
    stage = {"bow": 0.0, "match": 0.0, "mono": 0.0, "stereo": 0.0, "total": 0.0}
    acc = {"postings": 0, "hyp_m": 0, "hyp_s": 0, "pairs": 0, "res_m": 0, "res_s": 0}
    det.query_batch_upload(*batches[0])
    barrier()
    os.environ["KML_NO_GRAPH"] = "1"  # eager enqueue: a replayed graph carries no per-stage events
    for _ in range(KA):
        det.query_batch_run()
        st = det.stats()
        stage["bow"] += st.ms_bow; stage["match"] += st.ms_match
        stage["mono"] += st.ms_mono; stage["stereo"] += st.ms_stereo; stage["total"] += st.ms_total
        acc["postings"] += st.bow_postings_last; acc["hyp_m"] += st.mono_hypotheses_last
        acc["hyp_s"] += st.stereo_hypotheses_last; acc["pairs"] += st.pairs_last
        acc["res_m"] += st.mono_residuals_last; acc["res_s"] += st.stereo_residuals_last
    postings, hyp_m, hyp_s, pairs = acc["postings"], acc["hyp_m"], acc["hyp_s"], acc["pairs"]
    res_m, res_s = acc["res_m"], acc["res_s"]
    os.environ.pop("KML_NO_GRAPH", None)
    barrier()

    # ---------------- rooflines (algorithmic work / device time per stage)
    hbm_peak, peak_src, bf16_peak = load_peaks()
    traffic = load_traffic()
    compares = pairs * F * F
    roof = {
        "bow_scan": {"bound": "hbm", "achieved": postings * 8 / (stage["bow"] * 1e-3) / 1e9 if stage["bow"] else None,
                     "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                     "algorithmic_bytes_per_step": postings * 8 / KA, "ms_per_step": stage["bow"] / KA,
                     "traffic": traffic.get("bow_score_kernel_%s" % WORKLOAD.lower())},
        # algorithmic: 8 POPC32 per 256-bit compare; the kernel executes 5 (three carry-save adders fold
        # seven XOR words into two weight-1 and three weight-2 words first), so `achieved` can exceed
        # the pipe's peak; `frac` is the utilisation of the POPC pipe by what is actually executed
        "hamming_knn": ({"bound": "popc", "achieved": compares * 8 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                         "executed": compares * 5 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                         "peak": popc_peak / 1e12, "unit": "T POPC32/s", "peak_source": "measured (kml_peak_popc, same run)",
                         "compares_per_step": compares / KA, "ms_per_step": stage["match"] / KA, "engine": "popc",
                         "traffic": traffic.get("hamming_knn2_kernel_%s" % WORKLOAD.lower())}
                        if args.matcher_engine == 0 else
                        # tensor-core engine: one 256-bit compare = 256 int8 MACs = 512 ops on tcgen05.mma kind::i8;
                        # the same work expressed in POPC32 (8 per compare) is kept for comparison with the POPC pipe
                        {"bound": "tensor", "achieved": compares * 512 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                         "peak": 2.0 * bf16_peak, "unit": "TOP/s (int8)",
                         "peak_source": "2 x the measured dense bf16 rate of MEASURED_PEAKS.json (int8 tensor rate is nominally 2 x bf16)",
                         "compares_per_step": compares / KA, "ms_per_step": stage["match"] / KA, "engine": "tcgen05 kind::i8",
                         "equivalent_popc32_per_s": compares * 8 / (stage["match"] * 1e-3) / 1e12 if stage["match"] else None,
                         "popc_peak": popc_peak / 1e12,
                         "traffic": traffic.get("hamming_tc_kernel_%s" % WORKLOAD.lower())}),
        # algorithmic flops of the reference loop: per consumed hypothesis the minimal
        # solver (mono 33 kflop, stereo 1.5 kflop) plus one residual per correspondence (mono 95, stereo 27 flop)
        "mono_ransac": {"bound": "fp64", "hypotheses_per_step": hyp_m / KA, "residuals_per_step": res_m / KA,
                        "ms_per_step": stage["mono"] / KA,
                        "achieved": (hyp_m * 33e3 + res_m * 95.0) / (stage["mono"] * 1e-3) / 1e12 if stage["mono"] else None,
                        "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "peak_source": "measured (kml_peak_fp64, same run; DFMA — the "
                        "contract fuses where its oracle writes fma(), a fused multiply-add counted as 2 flop)", "traffic": None},
        "stereo_ransac": {"bound": "fp64", "hypotheses_per_step": hyp_s / KA, "residuals_per_step": res_s / KA,
                          "ms_per_step": stage["stereo"] / KA,
                          "achieved": (hyp_s * 1.5e3 + res_s * 27.0) / (stage["stereo"] * 1e-3) / 1e12 if stage["stereo"] else None,
                          "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "peak_source": "measured (kml_peak_fp64, same run)",
                          "traffic": None},
    }
    for k in roof:
        if roof[k]["achieved"]:
            roof[k]["frac"] = roof[k]["achieved"] / roof[k]["peak"]
    if roof["hamming_knn"].get("equivalent_popc32_per_s"):
        roof["hamming_knn"]["x_popc_pipe_peak"] = roof["hamming_knn"]["equivalent_popc32_per_s"] / roof["hamming_knn"]["popc_peak"]
    if roof["hamming_knn"].get("executed"):
        roof["hamming_knn"]["frac_algorithmic"] = roof["hamming_knn"]["frac"]
        roof["hamming_knn"]["frac"] = roof["hamming_knn"]["executed"] / roof["hamming_knn"]["peak"]
    dominant = max(("bow", "match", "mono", "stereo"), key=lambda s: stage[s])
    dom_key = {"bow": "bow_scan", "match": "hamming_knn", "mono": "mono_ransac", "stereo": "stereo_ransac"}[dominant]
    roofline = dict(roof[dom_key])
    roofline["kernel"] = dom_key
    roofline.setdefault("achieved", None)
    roofline.setdefault("frac", None)
    roofline.setdefault("traffic", None)

    valid = counts[:, None] > np.arange(out.shape[1])[None]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K,
        "warmup": args.warmup, "ms_per_step": step_ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64-fixed/f64", "data": "synthetic",
        "config": workload_config(world_size, sharded), "clocks": clocks, "gpu_launches": int(launches_blk),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_buffers": "pinned (kml_host_alloc)", "timed_blocks": len(e2e_ms)},
        "roofline": roofline, "rooflines": roof,
        "stage_ms_per_step": {k: v / KA for k, v in stage.items()},
        "stage_ms_source": "single-lane attribution pass of %d steps after the timed region" % KA,
        "timed_blocks": len(block_ms), "timed_region_s": float(np.sum(block_ms) * 1e-3),
        "block_ms_min_median_max": [float(np.min(block_ms)), step_ms, float(np.max(block_ms))],
        "lanes": n_lanes,
        "loop_closures_last_step": int((out["status"][valid] == 0).sum()),
    }
    if sharded:
        line["unit_note"] = ("value counts (query, shard) units: every one of the %d GPUs scores and verifies every query "
                             "against its own 50 000-keyframe shard" % world_size)
        line["distinct_queries_per_s"] = value / world_size
        line["sharded_parity"] = sharded_parity
        line["single_gpu_same_workload"] = single_same_workload
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(world, robots, log)
        fb = line["cpu_baseline"].get("fair_best")
        if fb:
            line["vs_fair_best"] = {"resident": value / fb["value"], "e2e": e2e_value / fb["value"],
                                    "note": "the -O3 -march=native oracle (hardware POPCNT), same host, same batch"}
    if rank == 0 and world_size == 1 and not args.no_configs:
        try:
            line["configs"] = side_configs(det, log, popc_peak, fp64_peak)
        except Exception as e:  # noqa: BLE001 - side measurements never take the bench line down
            line["configs"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    det.close()
    rdzv.close()


def _oracle_rate(ko, world, robots, threads, sample, log, tag):
    """queries/s of one `sample`-query batch on a freshly filled oracle detector of module `ko`"""
    lcd = ko.LoopClosureDetector()
    t0 = time.time()
    fill_oracle(lcd, world, robots, N_KEYFRAMES)
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
    try:  # the real OpenCV matcher on one candidate pair: 500 x 500, k = 2
        import cv2
        rng = np.random.default_rng(3)
        q = rng.integers(0, 256, (F, 32), np.uint8)
        t = rng.integers(0, 256, (F, 32), np.uint8)
        row = {}
        for name, nthr in (("1_thread_ms", 1), ("all_threads_ms", 0)):
            cv2.setNumThreads(nthr)
            m = cv2.BFMatcher(cv2.NORM_HAMMING)
            m.knnMatch(q, t, 2)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                m.knnMatch(q, t, 2)
                ts.append(time.perf_counter() - t0)
            row[name] = float(np.median(ts) * 1e3)
        row["what"] = "cv2.BFMatcher(NORM_HAMMING).knnMatch(500 x 500, k=2), median of 20, OpenCV %s" % cv2.__version__
        out["cv2_bfmatcher_pair"] = row
    except Exception as e:  # noqa: BLE001
        out["cv2_bfmatcher_pair"] = {"error": repr(e)}
    return out


def side_configs(det, log, popc_peak, fp64_peak):
    """BASELINE.json configs that are not the bench line, measured in the same driver run: C1 (one
    query against a 2 000-keyframe database: GPU latency through kml_query_batch with B = 1, oracle
    single-threaded), C3 (Hamming sweep 500 x 1k..1M against the measured POPC peak) and C4 (4 096
    stereo problems x 1 001 hypotheses x 500 correspondences, every hypothesis counted)."""
    import kml
    from kml import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kml_oracle as ko
    out = {}
    rng = np.random.default_rng(5)
    sweep = []
    for nt in (1000, 10000, 100000, 1000000):
        q = rng.integers(0, 256, (500, 32), np.uint8)
        t = rng.integers(0, 256, (nt, 32), np.uint8)
        _, _, ms = det.hamming_knn2(q, t, reps=10)
        row = {"nt": nt, "gpu_ms": ms, "gpu_compares_per_s": 500 * nt / (ms * 1e-3),
               "frac_of_popc_peak_algorithmic": 500 * nt * 8 / (ms * 1e-3) / popc_peak,
               "frac_of_popc_peak_executed": 500 * nt * 5 / (ms * 1e-3) / popc_peak}
        if nt <= 100000:
            try:
                import cv2
                cv2.setNumThreads(1)
                t0 = time.perf_counter()
                cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
                row["cv2_1thread_ms"] = (time.perf_counter() - t0) * 1e3
                cv2.setNumThreads(0)
            except Exception:  # noqa: BLE001
                pass
        sweep.append(row)
    out["C3_hamming_sweep"] = sweep
    # C4
    from scipy.spatial.transform import Rotation as Rot
    P, N = 4096, 500
    X = np.stack([rng.uniform(-5, 5, (P, N)), rng.uniform(-5, 5, (P, N)), rng.uniform(2, 12, (P, N))], axis=2)
    R = Rot.from_rotvec(rng.normal(size=(P, 3)) * 0.2).as_matrix()
    tt = rng.uniform(-1, 1, (P, 3))
    X2 = np.einsum("pnj,pjk->pnk", X - tt[:, None, :], R) + rng.normal(size=X.shape) * 0.03
    outl = rng.random((P, N)) < 0.35
    X2[outl] = rng.uniform(-8, 8, (int(outl.sum()), 3))
    det.ransac_arun_batch(X[:64], X2[:64], full_hypotheses=True)
    g = det.ransac_arun_batch(X, X2, full_hypotheses=True)
    hyp = float(P) * 1001
    res = hyp * N
    c4 = {"problems": P, "hypotheses_each": 1001, "correspondences": N, "gpu_ms": g["ms"],
          "residuals_per_s": res / (g["ms"] * 1e-3),
          "algorithmic_tflops": (hyp * 1.5e3 + res * 27) / (g["ms"] * 1e-3) / 1e12,
          "frac_of_fp64_peak": (hyp * 1.5e3 + res * 27) / (g["ms"] * 1e-3) / fp64_peak}
    t0 = time.perf_counter()
    for p in range(2):
        ko.ransac_arun(X[p], X2[p], 0.5, 1.0 - 1e-300, 1000, 12345)   # p ~ 1 keeps k large: all 1 001 trials
    c4["oracle_ms_per_problem_1thread"] = (time.perf_counter() - t0) / 2 * 1e3
    out["C4_stereo_ransac"] = c4
    # C1
    world = synth.World(500, F=500)
    ref = ko.LoopClosureDetector()
    d1 = kml.LoopClosureDetector(device=det.device)
    for ch in synth.build_database(world, [0], 2000, chunk=1000):
        d1.addBowVectors(ch["robot"], ch["poses"], ch["bow_off"], ch["bow_ids"], ch["bow_vals"])
        d1.addVLCFrames(ch["robot"], ch["poses"], ch["desc"], ch["bearings"], ch["points"])
    fill_oracle(ref, world, [0], 2000)
    q = synth.make_queries(world, 8, 2000, 1)
    fq, fp = q["frames"], q["prev"]
    gpu_t, cpu_t, same = [], [], True
    for b in range(8):
        o0, o1 = fq["bow_off"][b], fq["bow_off"][b + 1]
        p0, p1 = fp["bow_off"][b], fp["bow_off"][b + 1]
        a = (q["q_robot"][b:b + 1] + 7, q["q_pose"][b:b + 1], np.array([0, o1 - o0]), fq["bow_ids"][o0:o1], fq["bow_vals"][o0:o1],
             np.array([0, p1 - p0]), fp["bow_ids"][p0:p1], fp["bow_vals"][p0:p1], fq["desc"][b:b + 1], fq["bearings"][b:b + 1],
             fq["points"][b:b + 1])
        t0 = time.perf_counter()
        r1, c1 = d1.query_batch(*a)
        gpu_t.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        r0, c0 = ref.query_batch(*a, threads=1)
        cpu_t.append(time.perf_counter() - t0)
        same = same and np.array_equal(c0, c1) and np.array_equal(r0["mono_inliers"], r1["mono_inliers"])
    out["C1_single_query"] = {"gpu_ms_per_query_median": float(np.median(gpu_t[1:]) * 1e3),
                              "cpu_oracle_1thread_ms_per_query_median": float(np.median(cpu_t[1:]) * 1e3),
                              "candidates_verified": int(c1[0]), "records_match_oracle": bool(same)}
    d1.close()
    return out


if __name__ == "__main__":
    main()
