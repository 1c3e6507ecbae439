#!/usr/bin/env python
"""Benchmark of the hot path: training words/sec, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                  # Transformer-base 64x50 (BASELINE.json configs[1])
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...                           # the CPU arm (see below)
    python bench.py --model s2s-deep-gru | transformer-big         # configs[2] / configs[4]
    python bench.py --scaling strong                               # N > 1: the GLOBAL batch is split ceil(B/N) per rank
    python bench.py --padded                                       # lengths uniform in [T/2, T] (mask path)

A "step" = one full update of the reference's training loop body
(src/training/graph_group_singleton.cu:21-66 / graph_group_sync.cu:42-188):
build/forward/backward over one batch, gradient exchange (N > 1), gradient-norm
clipping and Adam.  N > 1 defaults to WEAK scaling (every rank trains on its own
batch of the configured size); --scaling strong splits ONE global batch over
the ranks as the reference's CorpusBatch::split does (src/data/corpus.h:145-169).

JSON line (rank 0):
  value      src+trg words/s, device-timed (CUDA events on the engine stream), batches
             pre-staged; the per-step index/mask upload is part of the captured step
  e2e        same metric through the C ABI with HOST batches handed over every step
             (mrn_trainer_set_batch) and the cost read back (blocking) every step
  parity     first-step cost and a logits slab of THIS arithmetic mode against the CPU oracle
             on the identical batch and initialisation (the north-star 1e-4 bar applies to
             the exact modes 0 and 2; the throughput modes report their error here)
  roofline   tcgen05 GEMM launches of the step (the only tensor-bound kernel family):
             algorithmic 2*M*N*K flops / in-graph kernel spans
  cpu_baseline  the CPU oracle (oracle/, kind "port": the reference has no CPU backend)
             timed on this box's host cores on full steps of the same workload
--impl reference: the same oracle as a separate arm (the reference itself is CUDA-only and
cannot be built without Boost/cuBLAS-era toolchains).  Both CPU legs run the
oracle in a CHILD PROCESS (clean OpenMP environment, wall-clock guard, one progress line per
step) on the same configuration as the GPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

VOCAB = 32000
MODE_NAMES = {0: "f32", 1: "bf16-packed", 2: "bf16x3", 3: "tf32", 4: "bf16"}


_SHAPE_OVERRIDE = {"batch": None, "len": None}  # --batch / --len: another batch shape of the same model (e.g. one rank's share of config E)


def model_config(name, pkg, gemm_mode):
    """BASELINE.json configs -> (trainer options, sentences, length, label)."""
    o, b, l, label = _model_config(name, pkg, gemm_mode)
    return o, _SHAPE_OVERRIDE["batch"] or b, _SHAPE_OVERRIDE["len"] or l, label


def _model_config(name, pkg, gemm_mode):
    if name == "transformer-base":
        return pkg.transformer_base_options(gemm_mode=gemm_mode), 64, 50, "Transformer-base (6+6, d=512, 8 heads, ffn 2048, V=32000)"
    if name == "transformer-big":
        o = pkg.transformer_base_options(gemm_mode=gemm_mode, workspace=16384)
        o.update({"dim-emb": 1024, "transformer-heads": 16, "transformer-dim-ffn": 4096})
        return o, 256, 80, "Transformer-big (6+6, d=1024, 16 heads, ffn 4096, V=32000)"
    if name == "s2s-deep-gru":
        o = {"type": "s2s", "dim-vocabs": [VOCAB, VOCAB], "dim-emb": 512, "dim-rnn": 1024, "enc-depth": 4, "dec-depth": 4,
             "enc-cell": "gru", "dec-cell": "gru", "cost-type": "ce-mean", "label-smoothing": 0, "optimizer": "adam",
             "learn-rate": 0.0001, "clip-norm": 1, "seed": 1234, "workspace": 16384, "gemm-mode": gemm_mode}
        return o, 64, 50, "deep GRU s2s (4+4, dim-emb 512, dim-rnn 1024, V=32000) with attention"
    raise SystemExit("unknown --model " + name)


def metric_name(model):
    return "source+target words/sec (device-timed) " + {"transformer-base": "Transformer-base", "transformer-big": "Transformer-big",
                                                        "s2s-deep-gru": "deep-GRU s2s"}[model]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def usable_cores():
    """Host threads this process may really use: the affinity mask AND the cgroup CPU quota of the
    lease (os.cpu_count() reports the machine, not the lease)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as fh:
                txt = fh.read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fh2:
                        n = min(n, max(1, q // int(fh2.read().strip())))
            break
        except Exception:
            continue
    return max(1, n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_host_batches(n, seed, batch, length):
    """Dense batches in the reference's SubBatch layout (time-major [T, B])."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        src = rs.randint(2, VOCAB, size=(length, batch)).astype(np.int64)
        trg = rs.randint(2, VOCAB, size=(length, batch)).astype(np.int64)
        src[-1] = 0  # EOS
        trg[-1] = 0
        out.append((src, np.ones((length, batch), np.float32), trg, np.ones((length, batch), np.float32)))
    return out


# =============================================================================
# CPU oracle in a child process (cpu_baseline, --impl reference, parity vectors)
# =============================================================================
def oracle_child(spec_path):
    """Runs in the CHILD: one line of JSON per finished step on stdout (flushed)."""
    with open(spec_path) as fh:
        spec = json.load(fh)
    oracle = graft.load_oracle()
    t = oracle.trainer(spec["options"])
    B, L = spec["batch"], spec["length"]
    for i in range(spec["max_steps"]):
        t0 = time.perf_counter()
        t.next_synthetic_batch(B, L, L, padded=spec["padded"])
        keep = i == 0 and spec.get("dump")
        t.compute_gradients(keep_logits=bool(keep))
        rec = {"step": i}
        if keep:
            rec["cost0"] = t.cost()
            logits = t.get_tensor("logits").reshape(-1, spec["vocab"])
            np.save(spec["dump"], logits[::spec["row_stride"]])
            del logits
        t.update()
        rec["cost"] = t.cost()
        rec["sec"] = time.perf_counter() - t0
        rec["words"] = t.batch_words()[1]
        print(json.dumps(rec), flush=True)
    t.close()


def run_oracle(options, batch, length, padded, max_steps, budget_s, dump=None, row_stride=8, log=None):
    """Parent side: launches the child with a clean OpenMP environment, reads its per-step lines
    until `max_steps` are done or `budget_s` of wall-clock has passed, then stops it.  Returns
    (records, threads)."""
    threads = min(usable_cores(), 64)
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(threads), "OMP_WAIT_POLICY": "PASSIVE", "OMP_PROC_BIND": "false", "OMP_DYNAMIC": "false"})
    if isinstance(options, dict):
        options = ";".join("%s=%s" % (k, ",".join(map(str, v)) if isinstance(v, (list, tuple)) else v) for k, v in options.items())
    spec = {"options": options, "batch": batch, "length": length, "padded": bool(padded), "max_steps": max_steps,
            "dump": dump, "row_stride": row_stride, "vocab": VOCAB}
    fd, spec_path = tempfile.mkstemp(suffix=".json")
    with os.fdopen(fd, "w") as fh:
        json.dump(spec, fh)
    proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--oracle-child", spec_path], stdout=subprocess.PIPE, text=True, env=env)
    recs = []

    def reader():
        for line in proc.stdout:
            line = line.strip()
            if line.startswith("{"):
                recs.append(json.loads(line))
                if log:
                    log("[cpu oracle] step %d: %.2f s, cost %.4f (%d threads)" % (recs[-1]["step"], recs[-1]["sec"], recs[-1]["cost"], threads))

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    deadline = time.time() + budget_s
    while th.is_alive() and time.time() < deadline:
        th.join(timeout=0.5)
    if proc.poll() is None:
        proc.kill()
    proc.wait()
    th.join(timeout=2)
    try:
        os.unlink(spec_path)
    except OSError:
        pass
    return recs, threads


def cpu_numbers(recs, threads, what):
    """words/s over the steps after the first (page-faulting) one; the first alone if nothing else finished."""
    timed = recs[1:] if len(recs) > 1 else recs
    if not timed:
        return None
    sec = sum(r["sec"] for r in timed)
    words = sum(r["words"] for r in timed)
    return {"value": words / sec, "unit": "words/s", "cores": threads, "kind": "port", "sec_per_step": sec / len(timed),
            "sample": "%d full step(s) of %s after %d warm-up step(s) of the same shape; CPU oracle (the reference has no CPU backend), OpenMP, %d threads"
                      % (len(timed), what, len(recs) - len(timed), threads)}


def run_reference_arm(args, rank):
    """CPU arm: the oracle port of the reference's hot path on the host cores, same configuration."""
    if rank != 0:
        return
    pkg = graft.load_package()
    opts, B, L, label = model_config(args.model, pkg, 0)
    opts["graph-replay"] = "false"
    what = "%s, dense %dx%d" % (label, B, L)
    steps = max(1, min(args.steps, 5))
    recs, threads = run_oracle(opts, B, L, args.padded, 1 + steps, float(os.environ.get("MRN_REFERENCE_BUDGET_S", "110")), log=lambda s: print(s, file=sys.stderr, flush=True))
    num = cpu_numbers(recs, threads, what)
    if num is None:
        print(json.dumps({"impl": "reference", "unavailable": "the CPU oracle finished no step inside the wall budget (%d threads)" % threads}))
        return
    n_timed = len(recs) - 1 if len(recs) > 1 else 1
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.model), "value": num["value"], "unit": "words/s", "n_gpus": args.gpus, "steps": n_timed,
        "steps_requested": args.steps, "warmup": len(recs) - n_timed,
        "ms_per_step": 1000 * num["sec_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s training step, %s %dx%d-token bitext per GPU" % (label, "padded" if args.padded else "dense", B, L),
                   "global_batch": B, "seq_len": L,
                   "note": "CPU arm: rank 0 only, one batch of the per-GPU size per step; the number of timed steps is capped by a wall budget"},
        "cpu_baseline": num,
        "e2e": {"value": num["value"], "unit": "words/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# =============================================================================
# the GPU arm
# =============================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="transformer-base", choices=["transformer-base", "transformer-big", "s2s-deep-gru"])
    ap.add_argument("--batch", type=int, default=None, help="sentences per batch instead of the config's own (a rank's share of a global batch)")
    ap.add_argument("--len", type=int, default=None, help="tokens per sentence instead of the config's own")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--padded", action="store_true", help="sentence lengths uniform in [T/2, T] (mask path) instead of dense batches")
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("MRN_BENCH_GEMM_MODE", "4")),
                    help="3 = tf32 on the fp32 tensors, 4 = bf16 operands (shadow copies), 1 = packed bf16, 2 = bf16x3, 0 = fp32 SIMT")
    ap.add_argument("--no-graph-replay", action="store_true", help="eager tape every step (profiling aid: every kernel is an ordinary launch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange by peer-memory kernels over NVLink (default) or NCCL reduce-scatter / all-gather")
    ap.add_argument("--oracle-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu pass that measures the GEMM family's DRAM traffic")
    args = ap.parse_args()
    _SHAPE_OVERRIDE["batch"], _SHAPE_OVERRIDE["len"] = args.batch, args.len

    if args.oracle_child:
        oracle_child(args.oracle_child)
        return
    if args.traffic_child:
        traffic_child(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    pkg = graft.load_package()
    lib = pkg.load()
    lib.call("mrn_set_device", local_rank)
    # engine work, NCCL collectives and the timing events all live on ONE side stream
    # (the legacy default stream cannot be graph-captured)
    side = torch.cuda.Stream(priority=-1)  # the engine's chain stream: above its side streams (tensors/device_gpu.cu)
    torch.cuda.set_stream(side)
    lib.set_stream(side.cuda_stream)

    W = max(3, args.warmup)
    K = args.steps
    opts, BATCH, LEN, label = model_config(args.model, pkg, args.gemm_mode)
    if args.no_graph_replay:
        opts["graph-replay"] = "false"
    strong = args.scaling == "strong" and world > 1
    # strong scaling: every rank draws the SAME global batch and keeps split(N)[rank]
    opts["data-seed"] = 1111 if strong else 1111 + rank
    split = (rank, world) if strong else (0, 1)

    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # gradient exchange: native peer-memory kernels (default) or NCCL collectives (--exchange nccl)
        sync = pkg.SyncTrainer(lib, opts, local_rank, rank, world, pkg.TorchExchange(), peer=(args.exchange == "peer"))
        trainer = sync.trainer
        step = sync.step
        barrier = dist.barrier
    else:
        trainer = lib.trainer(opts, device=local_rank)

        def step():
            trainer.compute_gradients()
            trainer.update()

        def barrier():
            pass

    words_seen = [0]

    def timed(fn_step, steps):
        barrier()
        torch.cuda.synchronize()
        words_seen[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn_step(i)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        words = float(words_seen[0])
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
            w = torch.tensor([words], device="cuda", dtype=torch.float64)
            torch.distributed.all_reduce(w)
            words = float(w.item())
        return ms, words

    # ---- device-resident arm ("value"): batches are generated by the library's synthetic
    #      corpus outside the device's critical path; a step = staged index upload + replay ----
    def dev_step(i):
        trainer.next_synthetic_batch(BATCH, LEN, LEN, padded=args.padded, split_rank=split[0], split_n=split[1])
        words_seen[0] += trainer.batch_words()[1]
        step()

    for i in range(W):
        dev_step(i)
    cost_warm = trainer.cost()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, words = timed(dev_step, K)
    clocks = sampler.stop() if rank == 0 else None
    cost_end = trainer.cost()
    value = words / (ms / 1000.0)

    # ---- end-to-end arm: host batches through the C ABI every step + blocking cost read ----
    if strong:
        lo = rank * ((BATCH + world - 1) // world)
        hi = min(BATCH, lo + (BATCH + world - 1) // world)
        host = [tuple(a[:, lo:hi] for a in hb) for hb in synthetic_host_batches(8, 4242, BATCH, LEN)]
    else:
        host = synthetic_host_batches(8, 4242 + rank, BATCH, LEN)
    local_sent = host[0][0].shape[1]

    def e2e_step(i):
        s, sm, t, tm = host[i % len(host)]
        trainer.set_batch(s, sm, t, tm)
        words_seen[0] += int(sm.sum() + tm.sum())
        step()
        trainer.cost()  # blocking 4-byte D2H

    for i in range(2):
        e2e_step(i)
    ms_e2e, words_e2e = timed(e2e_step, K)
    e2e_value = words_e2e / (ms_e2e / 1000.0)
    # per step and rank: src/trg indices as int32 rows, src/trg masks and the float labels
    h2d = 4 * (2 * local_sent * LEN) + 4 * (2 * local_sent * LEN) + 4 * (local_sent * LEN)
    if opts["type"] == "transformer":
        # + the shape-only constants re-uploaded by the captured step (positional signal 2x, triangle mask)
        h2d += 4 * (2 * LEN * int(opts["dim-emb"]) + LEN * LEN)

    stats = trainer.stats()
    graph_kernels = trainer.graph_kernels()
    # + sum-of-squares, Adam (+ NCCL RS/AG), or with the peer exchange: 2 barriers, gather-reduce, Adam
    launches = graph_kernels + (2 if world == 1 else 4)

    if rank == 0:
        pk, pk_src = peaks()
        prof = spans = None
        if world == 1:
            # roofline of the tensor-core GEMM family: algorithmic flops of one step / their device time
            prof = gemm_profile(lib, opts, BATCH, LEN, args.padded, local_rank)
            if prof:
                os.environ["MRN_GEMM_SPANS"] = "1"  # per-launch kernel execution spans from %globaltimer, no event nodes
                try:
                    spans = gemm_profile(lib, opts, BATCH, LEN, args.padded, local_rank)
                finally:
                    os.environ.pop("MRN_GEMM_SPANS", None)
        shape = "%s %dx%d-token bitext" % ("padded" if args.padded else "dense", BATCH, LEN)
        out = {
            "metric": metric_name(args.model), "value": value, "unit": "words/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": MODE_NAMES[args.gemm_mode], "data": "synthetic",
            "config": {"workload": "%s training step, %s %s" % (label, shape, "split over the GPUs" if strong else "per GPU"),
                       "global_batch": BATCH if strong else world * BATCH, "seq_len": LEN, "parallelism": "dp%d" % world,
                       "exchange": (("peer-memory kernels over NVLink" if sync.peer else "NCCL reduce-scatter / all-gather") if world > 1 else None),
                       "l2": "working set per step (parameters + gradients + activations) >> 126 MB L2; no explicit flush",
                       "gemm_mode": args.gemm_mode, "graph_replay": stats["plans"] > 0},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "words/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches * K,
            "gpu_launches_per_step": launches,
            "cost_first_last": [cost_warm, cost_end],
            "source_words_per_s": value / 2 if not args.padded else None,
        }
        if prof:
            main_ = spans if spans else prof
            out["roofline"] = {"bound": "tensor", "achieved": main_["tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": main_["tflops"] / pk["bf16_tflops_sustained"],
                               "traffic": None, "traffic_unit": "DRAM bytes per launch (mean over the step's GEMM launches)",
                               "algorithmic_flop_per_launch": prof["gflop"] * 1e9 / prof["launches"], "peak_source": pk_src,
                               "kernel": "tcgen05 GEMM family (all Prod/ProdBatched/ProdAffine launches of one replayed step)",
                               "duration_source": "in-kernel %globaltimer spans" if spans else "CUDA event pairs inside the graph",
                               "launches_per_step": prof["launches"], "gemm_ms_per_step": main_["ms"], "gflop_per_step": prof["gflop"],
                               "cuda_event_pairs": {"gemm_ms_per_step": prof["ms"], "achieved": prof["tflops"], "frac": prof["tflops"] / pk["bf16_tflops_sustained"],
                                                    "note": "each pair includes ~6.7 us of event-record node latency"}}
        if prof and not args.no_traffic:
            traffic, detail = measure_gemm_traffic(args, prof["launches"])
            out["roofline"]["traffic"] = traffic
            out["roofline"]["traffic_detail"] = detail
        # ---- CPU oracle (child process): parity vectors of step 1 + the timed CPU baseline ----
        if world == 1 and not (args.no_cpu_baseline and args.no_parity):
            dump = os.path.join(tempfile.gettempdir(), "mrn_oracle_logits_%d.npy" % os.getpid())
            o_opts = dict(opts)
            o_opts["gemm-mode"] = 0
            o_opts["graph-replay"] = "false"
            o_opts["data-seed"] = 1111
            budget = float(os.environ.get("MRN_CPU_BASELINE_BUDGET_S", "100"))
            recs, threads = run_oracle(o_opts, BATCH, LEN, args.padded, 1 if args.no_cpu_baseline else 4, budget,
                                       dump=None if args.no_parity else dump, log=lambda s: print(s, file=sys.stderr, flush=True))
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_numbers(recs, threads, "%s, %s" % (label, shape))
            if not args.no_parity and recs and "cost0" in recs[0] and os.path.exists(dump):
                out["parity"] = parity_vs_oracle(lib, opts, BATCH, LEN, args.padded, local_rank, recs[0]["cost0"], np.load(dump), 8, args.gemm_mode)
            try:
                os.unlink(dump)
            except OSError:
                pass
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def traffic_child(args):
    """Runs under ncu (measure_gemm_traffic): two EAGER steps of the benchmarked configuration - every GEMM is an
    ordinary launch; the parent takes the launches of the second step."""
    pkg = graft.load_package()
    lib = pkg.load()
    opts, B, L, _ = model_config(args.model, pkg, args.gemm_mode)
    opts["graph-replay"] = "false"
    t = lib.trainer(opts)
    for _ in range(2):
        t.next_synthetic_batch(B, L, L, padded=False)
        t.compute_gradients()
        t.update()
        t.cost()
    t.close()


def measure_gemm_traffic(args, launches_per_step):
    """DRAM bytes of the GEMM family, measured in THIS run: one `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
    pass over the GEMM kernels of an eager step in a child process.  Returns (mean bytes per launch, detail) or
    (None, reason) when ncu is unavailable / fails / takes too long."""
    import csv
    import shutil

    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if not ncu:
        return None, "ncu not found"
    log = os.path.join(tempfile.gettempdir(), "mrn_gemm_traffic_%d.csv" % os.getpid())
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:gGemm", "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--model", args.model, "--gemm-mode", str(args.gemm_mode)]
    if args.batch:
        cmd += ["--batch", str(args.batch)]
    if args.len:
        cmd += ["--len", str(args.len)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=float(os.environ.get("MRN_TRAFFIC_BUDGET_S", "150")))
    except subprocess.TimeoutExpired:
        return None, "ncu pass exceeded its wall budget"
    if r.returncode != 0 or not os.path.exists(log):
        return None, "ncu pass failed (rc %d)" % r.returncode
    per_id = {}
    with open(log, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row.get("Metric Unit", "byte"), 1)
        per_id.setdefault(int(row["ID"]), [0.0, 0.0])[0 if row["Metric Name"].endswith("read.sum") else 1] += v
    try:
        os.unlink(log)
    except OSError:
        pass
    ids = sorted(per_id)
    if len(ids) < 2:
        return None, "ncu pass saw no GEMM launches"
    step = ids[len(ids) // 2:]  # second of the two eager steps
    rd, wr = sum(per_id[i][0] for i in step), sum(per_id[i][1] for i in step)
    return (rd + wr) / len(step), {"launches": len(step), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
                                    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gGemm over an eager step, measured in this run"}


def parity_vs_oracle(lib, opts, B, L, padded, device, cost_ref, logits_ref, row_stride, mode):
    """First step of a fresh trainer (same seed -> same initialisation, same first synthetic batch)
    in the benchmarked arithmetic mode against the oracle's cost and logits slab."""
    o = dict(opts)
    o["data-seed"] = 1111
    o["graph-replay"] = "false"
    t = lib.trainer(o, device=device)
    t.next_synthetic_batch(B, L, L, padded=padded)
    t.compute_gradients(keep_logits=True)
    cost = t.cost()
    logits = t.get_tensor("logits").reshape(-1, VOCAB)[::row_stride]
    t.close()
    scale = max(1e-6, float(np.abs(logits_ref).max()))
    err = float(np.abs(logits.astype(np.float64) - logits_ref).max()) / scale
    tol = {0: 1e-4, 2: 1e-4, 3: 3e-3}.get(mode, 2e-2)  # the tolerances of tests/test_gpu_fullsize.py
    rel = abs(cost - cost_ref) / abs(cost_ref)
    return {"mode": MODE_NAMES[mode], "against": "CPU oracle, identical batch and initialisation, step 1",
            "tolerance": tol, "ok": bool(rel <= tol and err <= tol),
            "cost": cost, "cost_oracle": cost_ref, "cost_rel_err": abs(cost - cost_ref) / abs(cost_ref),
            "logits_max_rel_err": err, "logits_rows_compared": int(logits_ref.shape[0]), "logits_row_stride": row_stride,
            "bar": "1e-4 for the exact modes (f32, bf16x3: tests/test_gpu_fullsize.py); throughput modes report their error"}


def gemm_profile(lib, opts, B, L, padded, device):
    """In-graph duration of every tensor-core GEMM launch of ONE replayed step: CUDA events are
    recorded on the engine stream around each launch while the step is captured (external
    event-record nodes, csrc/kernels/gemm.cu ProfileScope) and re-stamped by every replay - or,
    with MRN_GEMM_SPANS=1, the kernels fold their own %globaltimer start/end per launch.  This
    is a separate trainer: the probes perturb the step a little, so `value` is not taken here."""
    import ctypes

    fn = lib.c.mrn_gemm_profile
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)]
    fn.restype = ctypes.c_int
    t = lib.trainer(opts, device=device)

    def step():
        t.next_synthetic_batch(B, L, L, padded=False)  # one shape -> one plan (a padded run is profiled on the dense shape)
        t.compute_gradients()
        t.update()
        t.cost()

    step()  # eager: parameters, arenas
    ms, flops, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_size_t()
    fn(1, ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(n))  # enable + reset
    step()  # this shape's second appearance: captured with the probes, then launched
    for _ in range(4):
        step()  # replays re-stamp
    stats = t.stats()
    fn(0, ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(n))  # disable + read the last replay
    t.close()
    if n.value == 0 or ms.value <= 0 or stats["plans"] < 1:
        return None
    return {"tflops": flops.value / (ms.value / 1000.0) / 1e12, "ms": ms.value, "gflop": flops.value / 1e9, "launches": n.value}


if __name__ == "__main__":
    main()
