"""Host logic of AsyncGraphGroup (reference training/graph_group_async.cu:16-250) on the CPU
oracle: one rank fetching from / pushing into its own master shards through the lock + remote-Adam
statements (oracle/cpu/tensor_operators_cpu.cpp) must reproduce SingletonGraph; optimizer-delay
accumulates tau batches per push.  The multi-GPU form of the same code (peer memory over NVLink) is
covered by tests/test_gpu_model.py and scripts/async_check.py."""
import numpy as np

MODEL = ("type=transformer;dim-vocabs=60,70;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;enc-depth=1;dec-depth=1;"
         "workspace=128;gemm-mode=0;graph-replay=false")


def test_async_single_rank_equals_singleton(oracle, pkg):
    opts = MODEL + ";learn-rate=0.001"
    ref = oracle.trainer(opts)  # one trainer after the other: parameter initialisation draws from a process-wide seed counter
    ref_costs = []
    for s in range(4):
        ref.next_synthetic_batch(4, 7, 9, padded=True)
        ref.compute_gradients()
        ref.update()
        ref_costs.append(ref.cost())
    pr = ref.arena_numpy("params")
    ref.close()

    a = pkg.AsyncTrainer(oracle, opts, 0)
    costs = []
    for s in range(4):
        a.trainer.next_synthetic_batch(4, 7, 9, padded=True)
        a.step()
        costs.append(a.cost())
    a.fetch()  # replica <- master shards
    pa = a.trainer.arena_numpy("params")
    a.trainer.close()
    assert np.allclose(costs, ref_costs, rtol=1e-6), (costs, ref_costs)
    # Adam moves weights with analytically zero gradients (key biases) by rounding noise; all others agree
    diff = np.abs(pa - pr)
    assert np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 1e-6, (diff.max(), np.mean(diff > 2e-5))


def test_async_optimizer_delay_accumulates(oracle, pkg):
    """optimizer-delay tau = 2 (reference :186-215): parameters only move on every second step."""
    a = pkg.AsyncTrainer(oracle, MODEL + ";learn-rate=0.002;optimizer-delay=2", 0)
    snaps, costs = [], []
    for s in range(6):
        a.trainer.next_synthetic_batch(4, 7, 9, padded=False)
        a.step()
        costs.append(a.cost())
        a.fetch()
        snaps.append(a.trainer.arena_numpy("params").copy())
    a.trainer.close()
    assert all(np.isfinite(costs)), costs
    moved = [bool(np.any(snaps[i] != snaps[i - 1])) for i in range(1, 6)]
    # pushes happen after steps 2, 4, 6 (indices 1, 3, 5): the master parameters change exactly there
    assert moved == [True, False, True, False, True], moved
