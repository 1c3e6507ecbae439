#!/usr/bin/env python
"""Benchmark of the hot path: training words/sec of Transformer-base
(BASELINE.json configs[1]: 6+6 layers, d=512, 8 heads, ffn 2048, V=32000,
bf16 tensor-core GEMMs with fp32 accumulate/master weights) on synthetic dense
64 x 50-token bitext, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (see below)

A "step" = one full update of the reference's training loop body
(src/training/graph_group_singleton.cu:21-66 / graph_group_sync.cu:42-188):
build/forward/backward over one batch, gradient exchange (N > 1), gradient-norm
clipping and Adam.  N > 1 is WEAK scaling: every rank trains on its own 64 x 50
batch, gradients are averaged with reduce-scatter, the owned shard is updated
and parameters all-gathered (SyncGraphGroup semantics).

JSON line (rank 0):
  value      src+trg words/s, device-timed (CUDA events on the engine stream), batches
             pre-staged; the per-step 64 KB index/mask upload is part of the captured step
  e2e        same metric through the C ABI with HOST batches handed over every step
             (mrn_trainer_set_batch) and the cost read back (blocking) every step
  roofline   tcgen05 GEMM launches of the step (the only tensor-bound kernel family):
             algorithmic 2*M*N*K flops / CUDA-event time per launch, over eager steps
  cpu_baseline  the CPU oracle (oracle/, kind "port": the reference has no CPU backend)
             timed on this box's host cores on one full 64 x 50 step
--impl reference: the same oracle as a separate arm (the reference itself is
CUDA-only and cannot be built without Boost/cuBLAS-era toolchains).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

BATCH, LEN = 64, 50
VOCAB = 32000
WORDS_PER_BATCH = 2 * BATCH * LEN  # source + target, dense
METRIC = "source+target words/sec (device-timed) Transformer-base"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


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


def synthetic_host_batches(n, seed, pinned_tensors=None):
    """Dense 64 x 50 batches in the reference's SubBatch layout (time-major [T, B])."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        src = rs.randint(2, VOCAB, size=(LEN, BATCH)).astype(np.int64)
        trg = rs.randint(2, VOCAB, size=(LEN, BATCH)).astype(np.int64)
        src[-1] = 0  # EOS
        trg[-1] = 0
        out.append((src, np.ones((LEN, BATCH), np.float32), trg, np.ones((LEN, BATCH), np.float32)))
    return out


def run_reference_arm(args, rank):
    """CPU arm: the oracle port of the reference's hot path on the host cores."""
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    oracle = graft.load_oracle()
    pkg = graft.load_package()
    cores = os.cpu_count()
    steps, warm = args.steps, max(1, args.warmup)
    # bounded sample: a full 64 x 50 step costs ~10-15 s on 8 cores; shrink the sentence
    # count so the whole arm stays within a few minutes
    budget_s = 150.0
    sent = int(max(4, min(BATCH, BATCH * budget_s / ((steps + warm) * 13.0))))
    t = oracle.trainer(pkg.transformer_base_options(gemm_mode=0, workspace=8192))
    for _ in range(warm):
        t.next_synthetic_batch(sent, LEN, LEN)
        t.compute_gradients()
        t.update()
        t.cost()
    t0 = time.perf_counter()
    for _ in range(steps):
        t.next_synthetic_batch(sent, LEN, LEN)
        t.compute_gradients()
        t.update()
        t.cost()
    dt = time.perf_counter() - t0
    words = 2 * sent * LEN * steps
    v = words / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "words/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1000 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Transformer-base 6+6 d512 V32000, dense %dx%d sample of the 64x50 batch per step" % (sent, LEN)},
        "cpu_baseline": {"value": v, "unit": "words/s", "cores": cores, "kind": "port",
                         "sample": "%d steps of %d x %d tokens (reference has no CPU backend; oracle port, OpenMP on all host cores)" % (steps, sent, LEN)},
        "e2e": {"value": v, "unit": "words/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample():
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
    oracle = graft.load_oracle()
    pkg = graft.load_package()
    t = oracle.trainer(pkg.transformer_base_options(gemm_mode=0, workspace=8192))
    # warm-up with the SAME shape: the first step of a shape touches ~4 GB of fresh arena pages
    # (page faults would otherwise dominate the timed step)
    t.next_synthetic_batch(BATCH, LEN, LEN)
    t.compute_gradients()
    t.update()
    t.cost()
    steps = 2
    t0 = time.perf_counter()
    for _ in range(steps):
        t.next_synthetic_batch(BATCH, LEN, LEN)
        t.compute_gradients()
        t.update()
        t.cost()
    dt = (time.perf_counter() - t0) / steps
    t.close()
    return {"value": WORDS_PER_BATCH / dt, "unit": "words/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d full steps (64 x 50 src + 64 x 50 trg tokens) after one warm-up step of the same shape, %.1f s per step" % (steps, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm-mode", type=int, default=3, help="3 = tf32 on the fp32 tensors (headline), 1 = packed bf16, 2 = bf16x3, 0 = fp32 SIMT")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange by peer-memory kernels over NVLink (default) or NCCL reduce-scatter / all-gather")
    args = ap.parse_args()

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
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    lib.set_stream(side.cuda_stream)

    W = max(3, args.warmup)
    K = args.steps
    opts = pkg.transformer_base_options(gemm_mode=args.gemm_mode)
    opts["data-seed"] = 1111 + rank

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

    def timed(fn_step, steps):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn_step(i)
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident arm ("value"): batches are generated by the library's synthetic
    #      corpus outside the timed region; inside it a step = staged 64 KB upload + replay ----
    def dev_step(i):
        trainer.next_synthetic_batch(BATCH, LEN, LEN)  # host-side index generation only (tiny)
        step()

    for i in range(W):
        dev_step(i)
    cost_warm = trainer.cost()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(dev_step, K)
    clocks = sampler.stop() if rank == 0 else None
    cost_end = trainer.cost()
    value = world * WORDS_PER_BATCH * K / (ms / 1000.0)

    # ---- end-to-end arm: host batches through the C ABI every step + blocking cost read ----
    host = synthetic_host_batches(8, 4242 + rank)

    def e2e_step(i):
        s, sm, t, tm = host[i % len(host)]
        trainer.set_batch(s, sm, t, tm)
        step()
        trainer.cost()  # blocking 4-byte D2H

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, K)
    e2e_value = world * WORDS_PER_BATCH * K / (ms_e2e / 1000.0)
    # per step: src/trg indices as int32 rows, src/trg masks and the float labels
    h2d = 4 * (2 * BATCH * LEN) + 4 * (2 * BATCH * LEN) + 4 * (BATCH * LEN)
    # + the shape-only constants re-uploaded by the captured step (positional signal 2x, triangle mask)
    h2d += 4 * (2 * LEN * 512 + LEN * LEN)

    stats = trainer.stats()
    graph_kernels = trainer.graph_kernels()
    # + sum-of-squares, Adam (+ NCCL RS/AG), or with the peer exchange: 2 barriers, gather-reduce, Adam
    launches = graph_kernels + (2 if world == 1 else 4)

    if rank == 0:
        pk, pk_src = peaks()
        # roofline of the tensor-core GEMM family: algorithmic flops of one step / their device time
        prof = gemm_profile(lib, pkg, args.gemm_mode, local_rank) if world == 1 else None
        # same replayed step without event nodes: per-launch kernel execution spans from %globaltimer
        spans = None
        if world == 1 and prof:
            os.environ["MRN_GEMM_SPANS"] = "1"
            try:
                spans = gemm_profile(lib, pkg, args.gemm_mode, local_rank)
            finally:
                os.environ.pop("MRN_GEMM_SPANS", None)
        out = {
            "metric": METRIC, "value": value, "unit": "words/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16", 2: "bf16x3", 3: "tf32"}[args.gemm_mode], "data": "synthetic",
            "config": {"workload": "Transformer-base (6+6, d=512, 8 heads, ffn 2048, V=32000) training step, dense 64x50-token bitext per GPU",
                       "global_batch": world * BATCH, "seq_len": LEN, "parallelism": "dp%d" % world,
                       "exchange": (("peer-memory kernels over NVLink" if sync.peer else "NCCL reduce-scatter / all-gather") if world > 1 else None),
                       "l2": "working set per step (373 MB params + 373 MB grads + activations) >> 126 MB L2; no explicit flush",
                       "gemm_mode": args.gemm_mode, "graph_replay": stats["plans"] > 0},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "words/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches * K,
            "gpu_launches_per_step": launches,
            "cost_first_last": [cost_warm, cost_end],
            "source_words_per_s": value / 2,
        }
        if prof:
            # DRAM bytes of the same kernel family from the committed ncu pass (profiles/gemm_dram_rNN.json,
            # dram__bytes_read.sum + dram__bytes_write.sum over all GEMM launches of one step), per launch
            traffic = None
            try:
                import glob
                latest = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gemm_dram_r*.json")))[-1]
                dj = json.load(open(latest))
                traffic = (dj["dram_read_bytes_per_step"] + dj["dram_write_bytes_per_step"]) / dj["launches_per_step"]
            except Exception:
                pass
            # Primary duration = kernel execution span inside the replayed graph (first CTA start .. last CTA
            # end from %globaltimer, folded per launch by the kernel itself; no extra graph nodes).  CUDA event
            # pairs recorded around each launch inside the graph are reported next to it: an external
            # event-record node costs ~3.3 us on this GPU (profiles/graph_gap_probe_r01.json), i.e. the event
            # figure charges ~6.7 us of graph-node latency to every launch.
            main = spans if spans else prof
            out["roofline"] = {"bound": "tensor", "achieved": main["tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": main["tflops"] / pk["bf16_tflops_sustained"], "traffic": traffic,
                               "traffic_unit": "DRAM bytes per launch (mean over the step's launches)",
                               "algorithmic_flop_per_launch": prof["gflop"] * 1e9 / prof["launches"], "peak_source": pk_src,
                               "kernel": "gGemmTf32 (all Prod/ProdBatched/ProdAffine launches of one replayed step)",
                               "duration_source": "in-kernel %globaltimer spans" if spans else "CUDA event pairs inside the graph",
                               "launches_per_step": prof["launches"], "gemm_ms_per_step": main["ms"], "gflop_per_step": prof["gflop"],
                               "cuda_event_pairs": {"gemm_ms_per_step": prof["ms"], "achieved": prof["tflops"], "frac": prof["tflops"] / pk["bf16_tflops_sustained"],
                                                    "note": "each pair includes ~6.7 us of event-record node latency"}}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_sample()  # rank 0 at N=1 only
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def gemm_profile(lib, pkg, gemm_mode, device):
    """In-graph duration of every tensor-core GEMM launch of ONE replayed step: CUDA events are
    recorded on the engine stream around each launch while the step is captured (external
    event-record nodes, csrc/kernels/gemm.cu ProfileScope) and re-stamped by every replay.  This
    is a separate trainer: the event nodes perturb the step a little, so `value` is not taken here."""
    import ctypes

    fn = lib.c.mrn_gemm_profile
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)]
    fn.restype = ctypes.c_int
    t = lib.trainer(pkg.transformer_base_options(gemm_mode=gemm_mode), device=device)

    def step():
        t.next_synthetic_batch(BATCH, LEN, LEN)
        t.compute_gradients()
        t.update()
        t.cost()

    step()  # eager: parameters, arenas
    ms, flops, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_size_t()
    fn(1, ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(n))  # enable + reset
    step()  # this shape's second appearance: captured with the event nodes, then launched
    for _ in range(4):
        step()  # replays re-stamp the events
    stats = t.stats()
    fn(0, ctypes.byref(ms), ctypes.byref(flops), ctypes.byref(n))  # disable + read the last replay
    t.close()
    if n.value == 0 or ms.value <= 0 or stats["plans"] < 1:
        return None
    return {"tflops": flops.value / (ms.value / 1000.0) / 1e12, "ms": ms.value, "gflop": flops.value / 1e9, "launches": n.value}


if __name__ == "__main__":
    main()
