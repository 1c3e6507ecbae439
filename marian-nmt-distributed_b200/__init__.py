"""marian-nmt-distributed_b200: B200-native hot path of Marian NMT v1.2.1
(tneck/marian-nmt-distributed): define-by-run autodiff graph over hand-written
sm_100a tensor operators + data-parallel gradient exchange.

The product is the native library lib/libmarian_b200.so (CUDA kernels + C++
host engine + the C ABI of include/marian_b200.h).  This package is the thin
host-side mirror used by tests, bench.py and the multi-GPU harness.  There is
NO CPU fallback: load() fails loudly when the library is missing or cannot run.
"""
import os

from .capi import DECLARED_SYMBOLS, DeviceArray, Library, MarianError, MrnTensor, Trainer, null_tensor  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmarian_b200.so")
TESTS_LIB_PATH = os.path.join(HERE, "lib", "libmarian_b200_tests.so")

_lib = None


def load():
    """Loads the CUDA product library (built by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MarianError("CUDA extension %s is missing - run `python __graft_entry__.py build`" % LIB_PATH)
        lib = Library(LIB_PATH, TESTS_LIB_PATH)
        if lib.backend != "cuda":
            raise MarianError("libmarian_b200.so does not report the cuda backend")
        _lib = lib
    return _lib


def transformer_base_options(vocab=32000, gemm_mode=4, **extra):
    """BASELINE.json config[1]: Transformer-base (6+6, d=512, 8 heads, ffn 2048)."""
    o = {
        "type": "transformer", "dim-vocabs": [vocab, vocab], "dim-emb": 512, "enc-depth": 6, "dec-depth": 6,
        "transformer-heads": 8, "transformer-dim-ffn": 2048, "transformer-postprocess": "dan",
        "cost-type": "ce-mean", "label-smoothing": 0, "optimizer": "adam", "learn-rate": 0.0001, "clip-norm": 1,
        "seed": 1234, "workspace": 8192, "gemm-mode": gemm_mode,
    }
    o.update(extra)
    return o


from .sync import AsyncTrainer, SyncTrainer, TorchExchange  # noqa: E402,F401
