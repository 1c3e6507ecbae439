"""torchrun worker of tests/test_gpu_multi.py: SyncGraphGroup on N GPUs (one process per GPU)
against the CPU oracle's SyncGraphGroup on the SAME split batches.

Every rank runs, step by step on identical data,
  * the CUDA trainer with the PEER-MEMORY exchange (csrc/kernels/exchange.cu: barrier, gather-reduce
    by peer loads over NVLink, clip + Adam with peer stores, barrier),
  * the CUDA trainer with the NCCL exchange (reduce-scatter, shard update, all-gather),
  * the CPU oracle trainer with a gloo exchange (the reference semantics of
    src/training/graph_group_sync.cu:42-188: sum / N, per-shard clipping norm, Adam per shard),
and checks: identical replicas on all ranks after every step, costs equal to the oracle's within the
exact-mode tolerance, parameters equal up to Adam's sign noise on zero-gradient weights.
Prints one JSON line on rank 0; exit code 0 only if every check passed on every rank.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

OPTS = ("type=transformer;dim-vocabs=200,220;dim-emb=64;transformer-heads=4;transformer-dim-ffn=128;enc-depth=2;dec-depth=2;"
        "workspace=256;learn-rate=0.001;clip-norm=1;optimizer=adam;seed=1234;data-seed=1111")
B, LS, LT, STEPS = 16, 11, 13, 4


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")
    pkg = graft.load_package()
    lib = pkg.load()
    lib.call("mrn_set_device", local)
    oracle = graft.load_oracle()
    mode = int(os.environ.get("MRN_TEST_GEMM_MODE", "2"))

    def run(make):
        sync = make()
        costs, same = [], True
        for _ in range(STEPS):
            sync.trainer.next_synthetic_batch(B, LS, LT, padded=True, split_rank=rank, split_n=world)
            sync.step()
            costs.append(sync.cost())
            if sync.cuda:
                p = torch.from_numpy(sync.trainer.arena_numpy("params")).cuda()
                lo, hi = p.clone(), p.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                same = same and bool(torch.equal(lo, hi))
        params = sync.trainer.arena_numpy("params")
        peer = getattr(sync, "peer", False)
        sync.trainer.close()
        return {"costs": costs, "params": params, "replicas_identical": same, "peer": peer}

    exp = run(lambda: pkg.SyncTrainer(oracle, OPTS + ";gemm-mode=0;graph-replay=false", 0, rank, world, pkg.TorchExchange(group=gloo)))
    out = {}
    for name, peer in (("peer", True), ("nccl", False)):
        got = run(lambda: pkg.SyncTrainer(lib, OPTS + ";gemm-mode=%d;graph-replay=true" % mode, local, rank, world, pkg.TorchExchange(), peer=peer))
        n = min(len(got["params"]), len(exp["params"]))
        diff = np.abs(got["params"][:n] - exp["params"][:n])
        tol = 3e-4 if mode in (0, 2) else 3e-2
        res = {
            "used_peer_exchange": bool(got["peer"]),
            "replicas_identical": got["replicas_identical"],
            "costs": got["costs"], "costs_oracle": exp["costs"],
            "cost_rel_err": float(np.max(np.abs(np.array(got["costs"]) - np.array(exp["costs"])) / np.abs(exp["costs"]))),
            "param_diff_max": float(diff.max()), "param_diff_median": float(np.median(diff)), "param_frac_above_2e-5": float(np.mean(diff > 2e-5)),
        }
        # Adam moves every weight by ~lr per step whatever the gradient's size: weights with analytically
        # zero gradients follow rounding noise (bounded by steps * lr), all others must agree
        res["ok"] = bool(res["replicas_identical"] and res["cost_rel_err"] <= tol and res["param_diff_max"] <= 2 * STEPS * 1e-3 + 1e-5
                         and (mode not in (0, 2) or (res["param_frac_above_2e-5"] < 0.02 and res["param_diff_median"] < 5e-6)))
        if name == "peer":
            res["ok"] = res["ok"] and res["used_peer_exchange"]
        out[name] = res
    flags = torch.tensor([1 if all(v["ok"] for v in out.values()) else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(int(flags.item()))
    if rank == 0:
        print(json.dumps({"world": world, "gemm_mode": mode, "ok": ok, "exchanges": out}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
