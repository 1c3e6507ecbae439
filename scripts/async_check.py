"""N GPUs (torchrun): AsyncGraphGroup - every rank trains WITHOUT synchronising with the others
(fetch / push through per-shard device locks over peer memory).  Checks: costs finite and falling,
and after a final barrier + fetch every rank holds identical parameters (= the master shards)."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
pkg = graft.load_package(); lib = pkg.load(); lib.call("mrn_set_device", local)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); lib.set_stream(st.cuda_stream)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
full = len(sys.argv) > 1 and sys.argv[1] == "full"
if full:
    opts = dict(pkg.transformer_base_options(), **{"data-seed": 500 + rank})
    B, Ts, Tt, steps = 64, 50, 50, 30
else:
    opts = ("type=transformer;dim-vocabs=500,520;dim-emb=128;transformer-heads=4;transformer-dim-ffn=256;enc-depth=2;dec-depth=2;"
            "workspace=512;gemm-mode=3;learn-rate=0.0005;clip-norm=1;data-seed=%d" % (500 + rank))
    B, Ts, Tt, steps = 16, 20, 22, 40
a = pkg.AsyncTrainer(lib, opts, local, rank, world, dist)
costs = []
a.trainer.next_synthetic_batch(B, Ts, Tt, padded=False)
a.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(steps):
    a.trainer.next_synthetic_batch(B, Ts, Tt, padded=False)
    a.step()
    if i % 5 == 0:
        costs.append(a.cost())
    if rank == 1 and i % 7 == 0 and not full:
        time.sleep(0.01)  # a straggler: nobody waits for it
torch.cuda.synchronize(); dt = time.perf_counter() - t0
dist.barrier()
a.fetch(); torch.cuda.synchronize()
mine = torch.from_numpy(a.trainer.arena_numpy("params")).cuda()
ref = mine.clone(); dist.broadcast(ref, src=0)
same = bool((mine == ref).all().item())
ok = same and all(np.isfinite(costs)) and (full or costs[-1] < costs[0])
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
wps = torch.tensor([2.0 * B * Ts * steps / dt], device="cuda"); dist.all_reduce(wps)
if rank == 0:
    print(json.dumps({"world": world, "ok": bool(flag.item()), "replicas_identical_after_fetch": same, "costs_rank0": costs,
                      "ms_per_step_rank0": 1000 * dt / steps, "words_per_s_all_ranks": float(wps.item()), "full_size": full}))
a.close()  # drains this rank, barrier, then frees its master shard
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
