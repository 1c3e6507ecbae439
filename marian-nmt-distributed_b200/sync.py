"""Host-side mirror of SyncGraphGroup for ONE PROCESS PER GPU.

Reference: src/training/graph_group_sync.cu:42-188.  Each rank owns a Trainer
(ExpressionGraph + model + shard optimizer, all native) and this class drives
one update:

    compute_gradients()                      native forward+backward (graph replay)
    reduce_scatter(sum) flat gradient arena  torch.distributed (NCCL over NVLink / gloo in CPU tests)
    update_shard()                           native fused {x 1/N, shard-norm clip, Adam}
    all_gather flat parameter arena          torch.distributed

torch is used only as plumbing: the arenas are exposed zero-copy as torch
tensors (CUDA array interface / numpy) and the engine runs on torch's current
CUDA stream so that collectives and kernels are ordered without host syncs.
"""
import ctypes

import numpy as np


class _CudaView:
    """Zero-copy __cuda_array_interface__ view of engine device memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class TorchExchange:
    """reduce-scatter / all-gather / broadcast over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def reduce_scatter(self, shard_out, flat):
        if self.dist.get_backend(self.group) == "gloo":
            # gloo has no reduce_scatter: all-reduce the arena, keep the owned shard
            self.dist.all_reduce(flat, group=self.group)
            n = shard_out.numel()
            shard_out.copy_(flat[self.rank * n:(self.rank + 1) * n])
        else:
            self.dist.reduce_scatter_tensor(shard_out, flat, group=self.group)

    def all_gather(self, flat, shard_elems):
        mine = flat[self.rank * shard_elems:(self.rank + 1) * shard_elems]
        if self.dist.get_backend(self.group) == "gloo":
            parts = [flat[r * shard_elems:(r + 1) * shard_elems] for r in range(self.world)]
            self.dist.all_gather(parts, mine.clone(), group=self.group)
        else:
            self.dist.all_gather_into_tensor(flat, mine, group=self.group)

    def broadcast(self, flat):
        self.dist.broadcast(flat, src=0, group=self.group)

    def mean_cost(self, c, device):
        import torch

        t = torch.tensor([c], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, group=self.group)
        return float(t.item()) / self.world


class SyncTrainer:
    """peer=True (CUDA, ranks on one node, Adam): the update after backward is the native
    peer-memory exchange - {barrier, gather-reduce by peer loads over NVLink, clip + Adam with peer
    stores, barrier}, csrc/kernels/exchange.cu - instead of NCCL reduce-scatter / all-gather around
    the shard update.  torch.distributed then only carries the 64-byte IPC handles at start-up,
    the first-step parameter broadcast and the cost mean."""

    def __init__(self, lib, options, device, rank, nranks, exchange, peer=False):
        import torch

        self.torch = torch
        self.lib = lib
        self.rank, self.nranks = rank, nranks
        self.exchange = exchange
        self.cuda = lib.backend == "cuda"
        self.device = torch.device("cuda", device) if self.cuda else torch.device("cpu")
        if self.cuda:
            torch.cuda.set_device(device)
            # Engine work and torch's collectives share ONE non-default stream: kernels,
            # NCCL and CUDA-event timing are then ordered by the stream with no host
            # synchronisation.  (The legacy default stream cannot be graph-captured.)
            if torch.cuda.current_stream().cuda_stream == 0:
                self.stream = torch.cuda.Stream(device, priority=-1)  # chain stream above the engine's side streams
                torch.cuda.set_stream(self.stream)
            lib.set_stream(torch.cuda.current_stream().cuda_stream)
        self.trainer = lib.trainer(options, device=device, rank=rank, nranks=nranks)
        self.first = True
        self._views = None
        self.peer = bool(peer) and self.cuda and nranks > 1
        self._peers_mapped = False

    def _tensor(self, ptr, n):
        if self.cuda:
            return self.torch.as_tensor(_CudaView(ptr, n), device=self.device)
        buf = (ctypes.c_float * n).from_address(ptr)
        return self.torch.from_numpy(np.frombuffer(buf, dtype=np.float32))

    def _arenas(self):
        if self._views is None:
            p, n = self.trainer.params_arena()
            g, ng = self.trainer.grads_arena()
            s, ns = self.trainer.shard_grads_arena()
            assert n == ng == ns * self.nranks, (n, ng, ns)
            self._views = (self._tensor(p, n), self._tensor(g, ng), self._tensor(s, ns), ns)
        return self._views

    def _map_peers(self):
        """Exchanges the CUDA IPC handles of {params, grads, signal pad} and maps the peers.
        Every rank takes part in both collectives whatever happens locally; returns True only if
        ALL ranks mapped all peers."""
        dist, group = self.exchange.dist, self.exchange.group
        err = None
        try:
            mine = self.trainer.ipc_export()
        except Exception as e:  # noqa: BLE001
            mine, err = None, e
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, mine, group=group)
        if err is None and all(h is not None for h in gathered):
            try:
                self.trainer.ipc_import(b"".join(gathered), self.nranks)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer could not export its IPC handles")
        if err is not None:
            print("[marian_b200] rank %d: peer-memory exchange unavailable (%s); using collectives" % (self.rank, err), flush=True)
        flag = self.torch.tensor([0 if err else 1], dtype=self.torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self._peers_mapped = True
        return bool(int(flag.item()))

    def step(self):
        """One update on the trainer's current batch (set via self.trainer.*batch*)."""
        self.trainer.compute_gradients()
        params, grads, shard, ns = self._arenas()
        redo = self.first
        if self.first:
            # reference :46-53: replicas start from graph 0's parameters BEFORE any gradient is
            # taken.  Parameters only exist once the tape has been built and run, so the first
            # pass above serves as the initialisation pass: broadcast, then recompute the first
            # gradients at the common parameter point.
            self.exchange.broadcast(params)
            self.first = False
        if self.peer and not self._peers_mapped:
            # all ranks must agree: if any rank cannot map its peers (no P2P between the GPUs, IPC
            # disabled in the container) everybody falls back to the collective exchange - loudly.
            # (Mapped BEFORE the recomputation below: steps built with mapped peers are split so that
            # the gradient exchange of the upper shards overlaps the rest of the backward sweep.)
            self.peer = self._map_peers()
        if redo:
            self.trainer.compute_gradients()
        if self.peer:
            self.trainer.update_peer()
            return
        self.exchange.reduce_scatter(shard, grads)
        self.trainer.update_shard()
        self.exchange.all_gather(params, ns)

    def cost(self):
        return self.exchange.mean_cost(self.trainer.cost(), self.device)


class AsyncTrainer:
    """Host-side mirror of AsyncGraphGroup for one process per GPU (csrc/training/graph_group.h).

    Reference: src/training/graph_group_async.cu.  Every rank trains on its own stream of batches
    with NO synchronisation with the others: it fetches the parameter shards from their owners,
    runs forward/backward, and pushes its gradient slices into the owners' master shards, all by
    kernels over peer memory under per-shard device locks.  torch.distributed is only used to
    exchange the 64-byte IPC handles at start-up (nranks == 1 needs no process group)."""

    def __init__(self, lib, options, device, rank=0, nranks=1, dist=None, group=None):
        if isinstance(options, dict):
            options = dict(options, **{"graph-group": "async"})
        else:
            options = options + ";graph-group=async"
        self.rank, self.nranks = rank, nranks
        self.trainer = lib.trainer(options, device=device, rank=rank, nranks=nranks)
        self.dist, self.group = dist, group
        self._ready = False

    def _start(self):
        """Needs a current batch (parameters are created by building the tape once)."""
        self.trainer.async_init()
        mine = self.trainer.async_export()
        if self.nranks > 1:
            gathered = [None] * self.nranks
            self.dist.all_gather_object(gathered, mine, group=self.group)
            self.trainer.async_import(b"".join(gathered), self.nranks)
            self.dist.barrier(group=self.group)  # every master shard is seeded before anyone pushes
        else:
            self.trainer.async_import(mine, 1)
        self._ready = True

    def step(self):
        if not self._ready:
            self._start()
        self.trainer.async_update()

    def fetch(self):
        self.trainer.async_fetch()

    def cost(self):
        return self.trainer.cost()

    def close(self):
        """Peers write into this rank's master shard with no host involvement of this process: every rank drains
        its own queue and all ranks meet before any master block is freed."""
        self.trainer.cost()  # blocks until this rank's queued pushes / fetches have run
        if self.nranks > 1 and self._ready:
            self.dist.barrier(group=self.group)
        self.trainer.close()
