"""ctypes binding of the C ABI declared in include/marian_b200.h.

`Library(path)` wraps ONE shared object exporting the mrn_* symbols.  The
product library is loaded by `marian_nmt_distributed_b200.load()`; tests load
the CPU oracle (oracle/_build/libmarian_oracle.so) through the same class, which
is what lets a parity test run the identical call sequence on both.
"""
import ctypes
import os

import numpy as np

c_float_p = ctypes.POINTER(ctypes.c_float)


class MrnTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("rank", ctypes.c_int), ("shape", ctypes.c_int * 4)]


class MarianError(RuntimeError):
    pass


# name -> (argtypes); every function returns int status unless listed in _SPECIAL
_T, _TP, _I, _F, _V, _SZ = MrnTensor, ctypes.POINTER(MrnTensor), ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_SIGNATURES = {
    "mrn_set_device": [_I],
    "mrn_set_stream": [_V],
    "mrn_synchronize": [],
    "mrn_malloc": [ctypes.POINTER(_V), _SZ],
    "mrn_free": [_V],
    "mrn_memcpy_h2d": [_V, _V, _SZ],
    "mrn_memcpy_d2h": [_V, _V, _SZ],
    "mrn_memset_zero": [_V, _SZ],
    "mrn_gemm_create": [ctypes.POINTER(_V), _I],
    "mrn_gemm_destroy": [_V],
    "mrn_gemm_set_mode": [_V, _I],
    "mrn_gemm_debug_stamps": [_V],
    "mrn_gemm_profile": [_I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_SZ)],
    "mrn_prod": [_V, _T, _T, _T, _I, _I, _F, _F],
    "mrn_prod_batched": [_V, _T, _T, _T, _I, _I, _F, _F],
    "mrn_prod_grouped_nt": [_V, _T, _TP, _TP, _I, _F],
    "mrn_prod_swish_grad_nt": [_V, _T, _T, _T, _T, _F],
    "mrn_prod_grouped_nt_sums": [_V, _T, _TP, _TP, _I, _F, _TP],
    "mrn_prod_swish_grad_nt_sums": [_V, _T, _T, _T, _T, _F, _T],
    "mrn_prod_shared_a": [_V, _TP, _T, _TP, _TP, _I, _I, _F, ctypes.POINTER(_I)],
    "mrn_prod_affine": [_V, _T, _T, _T, _T],
    "mrn_element": [ctypes.c_char_p, _T, _TP, _I, _F],
    "mrn_add": [ctypes.c_char_p, _F, _T, _TP, _I, _F],
    "mrn_softmax": [_T, _T, _TP],
    "mrn_logsoftmax": [_T, _T],
    "mrn_softmax_grad": [_T, _T, _T],
    "mrn_logsoftmax_grad": [_T, _T, _T],
    "mrn_cross_entropy_pick": [_T, _T, _T],
    "mrn_cross_entropy_pick_backward": [_T, _T, _T, _T],
    "mrn_layer_norm": [_T, _T, _T, _TP, _F],
    "mrn_layer_norm_grad": [_T, _T, _TP, _T, _T, _T, _T, _TP, _F],
    "mrn_residual_layer_norm": [_T, _T, _T, _T, _T, _F],
    "mrn_residual_layer_norm_grad": [_T, _T, _T, _T, _T, _T, _T, _T, _T, _T, _F],
    "mrn_multi_head_attention": [_T, _T, _T, _T, _T, _TP, _I, _F, _I],
    "mrn_multi_head_attention_grad": [_T, _T, _T, _T, _T, _T, _T, _T, _T, _I, _F, _I],
    "mrn_att": [_T, _T, _T, _T],
    "mrn_att_back": [_T, _T, _T, _T, _T, _T, _T],
    "mrn_gru_fast_forward": [_T, _TP, _I, _I],
    "mrn_gru_fast_backward": [_TP, _TP, _I, _T, _I],
    "mrn_lstm_cell_forward": [_T, _TP, _I],
    "mrn_lstm_output_forward": [_T, _TP, _I],
    "mrn_lstm_cell_backward": [_TP, _TP, _I, _T],
    "mrn_lstm_output_backward": [_TP, _TP, _I, _T],
    "mrn_highway_forward": [_T, _T, _T, _T],
    "mrn_highway_backward": [_T, _T, _T, _T, _T, _T, _T],
    "mrn_transpose_nd": [_T, _T, ctypes.POINTER(_I)],
    "mrn_concatenate": [_T, _TP, _I, _I],
    "mrn_deconcatenate": [_TP, _I, _T, _I],
    "mrn_copy_rows": [_T, _T, _V, _SZ],
    "mrn_paste_rows": [_T, _T, _V, _SZ],
    "mrn_shift": [_T, _T, ctypes.POINTER(_I), _I],
    "mrn_l2norm": [_T, c_float_p],
    "mrn_adam_step": [_T, _T, _T, _T, _F, _F, _F, _F, _I, _F, _F],
    "mrn_sgd_step": [_T, _T, _F, _F, _F],
    "mrn_adagrad_step": [_T, _T, _T, _F, _F, _F, _F],
    "mrn_dropout": [_T, _F, ctypes.c_ulonglong],
    "mrn_trainer_create": [ctypes.POINTER(_V), ctypes.c_char_p, _I, _I, _I],
    "mrn_trainer_destroy": [_V],
    "mrn_trainer_set_batch": [_V, _I, _I, _V, _V, _I, _V, _V],
    "mrn_trainer_next_synthetic_batch": [_V, _I, _I, _I, _I, _I, _I],
    "mrn_trainer_open_corpus": [_V, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p],
    "mrn_trainer_next_corpus_batch": [_V, ctypes.POINTER(_I)],
    "mrn_trainer_get_batch": [_V, _I, _V, _V, _SZ, ctypes.POINTER(_I), ctypes.POINTER(_I)],
    "mrn_trainer_validate": [_V, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, c_float_p, c_float_p, ctypes.POINTER(_SZ), ctypes.POINTER(_SZ)],
    "mrn_trainer_translate": [_V, ctypes.c_char_p, _I, _I, _V, _V, _V, _V],
    "mrn_trainer_translate_file": [_V, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_SZ)],
    "mrn_nth_element_ranges": [_T, _V, _V, _I, _V, _V],
    "mrn_nth_element_logsoftmax": [_T, _V, _I, _I, _I, _I, _I, _V, _V],
    "mrn_trainer_compute_gradients": [_V, _I],
    "mrn_trainer_update": [_V],
    "mrn_trainer_update_shard": [_V],
    "mrn_trainer_ipc_export": [_V, ctypes.c_char_p, _SZ],
    "mrn_trainer_ipc_import": [_V, ctypes.c_char_p, _I],
    "mrn_trainer_update_peer": [_V],
    "mrn_trainer_async_init": [_V],
    "mrn_trainer_async_export": [_V, ctypes.c_char_p, _SZ],
    "mrn_trainer_async_import": [_V, ctypes.c_char_p, _I],
    "mrn_trainer_async_update": [_V],
    "mrn_trainer_async_fetch": [_V],
    "mrn_trainer_cost": [_V, c_float_p],
    "mrn_trainer_save": [_V, ctypes.c_char_p, _I],
    "mrn_trainer_load": [_V, ctypes.c_char_p, _I],
    "mrn_trainer_params": [_V, ctypes.POINTER(_V), ctypes.POINTER(_SZ)],
    "mrn_trainer_grads": [_V, ctypes.POINTER(_V), ctypes.POINTER(_SZ)],
    "mrn_trainer_shard_grads": [_V, ctypes.POINTER(_V), ctypes.POINTER(_SZ)],
    "mrn_trainer_get_tensor": [_V, ctypes.c_char_p, _I, _V, _SZ, ctypes.POINTER(_SZ)],
    "mrn_trainer_param_names": [_V, ctypes.c_char_p, _SZ, ctypes.POINTER(_SZ)],
    "mrn_trainer_batch_words": [_V, ctypes.POINTER(_SZ), ctypes.POINTER(_SZ)],
    "mrn_trainer_graph_kernels": [_V, ctypes.POINTER(_SZ)],
    "mrn_trainer_stats": [_V, ctypes.POINTER(_SZ), ctypes.POINTER(_SZ), ctypes.POINTER(_SZ), ctypes.POINTER(_SZ)],
}
# exported by the golden-vector driver (tests/cpp/graph_golden.cpp)
_TEST_SIGNATURES = {"mrn_test_golden": [ctypes.c_char_p, _V, _SZ, ctypes.POINTER(_SZ)]}

DECLARED_SYMBOLS = sorted(list(_SIGNATURES) + ["mrn_last_error", "mrn_backend_name"])


class DeviceArray:
    """A float32 (or int32) array in the library's device memory."""

    def __init__(self, lib, shape, dtype=np.float32):
        self.lib = lib
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        ptr = ctypes.c_void_p()
        lib._ck(lib.c.mrn_malloc(ctypes.byref(ptr), max(self.nbytes, 256)))
        self.ptr = ptr.value

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        self.lib._ck(self.lib.c.mrn_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes))
        return self

    def numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.lib._ck(self.lib.c.mrn_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes))
        return out

    def zero(self):
        self.lib._ck(self.lib.c.mrn_memset_zero(self.ptr, self.nbytes))
        self.lib.synchronize()
        return self

    def t(self, shape=None):
        """mrn_tensor view (optionally with another Marian shape of the same size)."""
        shape = self.shape if shape is None else tuple(shape)
        assert int(np.prod(shape)) * self.dtype.itemsize == self.nbytes
        assert 1 <= len(shape) <= 4
        m = MrnTensor()
        m.data = self.ptr
        m.rank = len(shape)
        for i, s in enumerate(shape):
            m.shape[i] = int(s)
        m._owner = self  # keeps the device buffer alive as long as the view exists
        return m

    def free(self):
        if self.ptr:
            self.lib.c.mrn_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def null_tensor():
    m = MrnTensor()
    m.data = None
    m.rank = 1
    m.shape[0] = 1
    return m


class Library:
    def __init__(self, path, tests_path=None):
        if not os.path.exists(path):
            raise MarianError("shared library not found: %s (run __graft_entry__.build())" % path)
        self.path = path
        self.c = ctypes.CDLL(path)
        self.c.mrn_last_error.restype = ctypes.c_char_p
        self.c.mrn_backend_name.restype = ctypes.c_char_p
        for name, args in _SIGNATURES.items():
            fn = getattr(self.c, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        self.tests = None
        tpath = tests_path
        if tpath is None and hasattr(self.c, "mrn_test_golden"):
            self.tests = self.c
        elif tpath and os.path.exists(tpath):
            self.tests = ctypes.CDLL(tpath)
        if self.tests is not None:
            for name, args in _TEST_SIGNATURES.items():
                fn = getattr(self.tests, name)
                fn.argtypes = args
                fn.restype = ctypes.c_int

    # -- helpers -----------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise MarianError(self.c.mrn_last_error().decode(errors="replace"))

    @property
    def backend(self):
        return self.c.mrn_backend_name().decode()

    def synchronize(self):
        self._ck(self.c.mrn_synchronize())

    def set_stream(self, stream_ptr):
        self._ck(self.c.mrn_set_stream(stream_ptr))

    def array(self, a, dtype=np.float32):
        a = np.asarray(a, dtype=dtype)
        return DeviceArray(self, a.shape, dtype).upload(a)

    def zeros(self, shape, dtype=np.float32):
        return DeviceArray(self, shape, dtype).zero()

    def call(self, name, *args):
        self._ck(getattr(self.c, name)(*args))

    @staticmethod
    def tensor_list(ts):
        arr = (MrnTensor * len(ts))()
        for i, t in enumerate(ts):
            arr[i] = t
        arr._owners = list(ts)  # element assignment copies the struct, not the python-side owner
        return arr

    def golden(self, case):
        assert self.tests is not None, "golden-vector driver not linked into this library"
        n = ctypes.c_size_t(0)
        rc = self.tests.mrn_test_golden(case.encode(), None, 0, ctypes.byref(n))
        if rc != 0:
            raise MarianError("mrn_test_golden(%s) failed with %d: %s" % (case, rc, self.c.mrn_last_error().decode()))
        out = np.zeros(n.value, dtype=np.float32)
        rc = self.tests.mrn_test_golden(case.encode(), out.ctypes.data, n.value, ctypes.byref(n))
        if rc != 0:
            raise MarianError("mrn_test_golden(%s) failed with %d" % (case, rc))
        return out

    def gemm(self, mode=0, device=0):
        return Gemm(self, mode, device)

    def nth_element_ranges(self, scores, range_first, cum_n):
        """(costs, keys): cum_n[i+1]-cum_n[i] best (value, flat index) pairs of every range, best first."""
        first = np.ascontiguousarray(range_first, dtype=np.int32)
        cum = np.ascontiguousarray(cum_n, dtype=np.int32)
        costs = np.zeros(int(cum[-1]), dtype=np.float32)
        keys = np.zeros(int(cum[-1]), dtype=np.uint32)
        self._ck(self.c.mrn_nth_element_ranges(scores, first.ctypes.data, cum.ctypes.data, len(first) - 1, costs.ctypes.data, keys.ctypes.data))
        return costs, keys

    def nth_element_logsoftmax(self, logits, prev_costs, dim_batch, beam, n, first=False, suppress_word=-1):
        """(costs, keys) of the n best continuations per sentence from raw logits [beam, 1, batch, V]."""
        prev = np.ascontiguousarray(prev_costs, dtype=np.float32)
        costs = np.zeros(dim_batch * n, dtype=np.float32)
        keys = np.zeros(dim_batch * n, dtype=np.uint32)
        self._ck(self.c.mrn_nth_element_logsoftmax(logits, prev.ctypes.data, dim_batch, beam, n, int(first), suppress_word, costs.ctypes.data, keys.ctypes.data))
        return costs, keys

    def trainer(self, options, device=0, rank=0, nranks=1):
        return Trainer(self, options, device, rank, nranks)


class Gemm:
    def __init__(self, lib, mode, device):
        self.lib = lib
        h = ctypes.c_void_p()
        lib._ck(lib.c.mrn_gemm_create(ctypes.byref(h), device))
        self.h = h
        lib._ck(lib.c.mrn_gemm_set_mode(h, mode))

    def set_mode(self, mode):
        self.lib._ck(self.lib.c.mrn_gemm_set_mode(self.h, mode))

    def __del__(self):
        try:
            if self.h:
                self.lib.c.mrn_gemm_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Trainer:
    """Training-step driver (reference: SingletonGraph / SyncGraphGroup)."""

    def __init__(self, lib, options, device=0, rank=0, nranks=1):
        self.lib = lib
        if isinstance(options, dict):
            options = ";".join("%s=%s" % (k, ",".join(map(str, v)) if isinstance(v, (list, tuple)) else v) for k, v in options.items())
        self.options = options
        self.rank, self.nranks = rank, nranks
        h = ctypes.c_void_p()
        lib._ck(lib.c.mrn_trainer_create(ctypes.byref(h), options.encode(), device, rank, nranks))
        self.h = h

    def close(self):
        if self.h:
            self.lib._ck(self.lib.c.mrn_trainer_destroy(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch(self, src_idx, src_mask, trg_idx, trg_mask):
        """Arrays are time-major [T, B] as in the reference's SubBatch."""
        src_idx = np.ascontiguousarray(src_idx, dtype=np.int64)
        trg_idx = np.ascontiguousarray(trg_idx, dtype=np.int64)
        src_mask = np.ascontiguousarray(src_mask, dtype=np.float32)
        trg_mask = np.ascontiguousarray(trg_mask, dtype=np.float32)
        Ts, B = src_idx.shape
        Tt, B2 = trg_idx.shape
        assert B == B2
        self.lib._ck(self.lib.c.mrn_trainer_set_batch(self.h, B, Ts, src_idx.ctypes.data, src_mask.ctypes.data, Tt, trg_idx.ctypes.data, trg_mask.ctypes.data))

    def next_synthetic_batch(self, batch_size, len_src, len_trg, padded=False, split_rank=0, split_n=1):
        self.lib._ck(self.lib.c.mrn_trainer_next_synthetic_batch(self.h, batch_size, len_src, len_trg, int(padded), split_rank, split_n))

    def open_corpus(self, src_path, trg_path, vocab_src=None, vocab_trg=None, options=""):
        enc = lambda x: None if x is None else str(x).encode()
        self.lib._ck(self.lib.c.mrn_trainer_open_corpus(self.h, enc(src_path), enc(trg_path), enc(vocab_src), enc(vocab_trg), options.encode()))

    def next_corpus_batch(self):
        """Makes the next mini-batch of the text corpus current; False at the end of an epoch."""
        has = ctypes.c_int()
        self.lib._ck(self.lib.c.mrn_trainer_next_corpus_batch(self.h, ctypes.byref(has)))
        return bool(has.value)

    def validate(self, src_path, trg_path, vocab_src=None, vocab_trg=None, options=""):
        """Cross-entropy validation on a held-out corpus: {"metric", "cost_sum", "sentences", "target_words"}."""
        enc = lambda x: None if x is None else str(x).encode()
        m, c = ctypes.c_float(), ctypes.c_float()
        n, w = ctypes.c_size_t(), ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_validate(self.h, enc(src_path), enc(trg_path), enc(vocab_src), enc(vocab_trg), options.encode(), ctypes.byref(m), ctypes.byref(c),
                                                     ctypes.byref(n), ctypes.byref(w)))
        return {"metric": m.value, "cost_sum": c.value, "sentences": n.value, "target_words": w.value}

    def translate(self, options="", n_best=1, max_len=None):
        """Beam search over the source side of the current batch.  Returns, per sentence, a list of up to n_best
        (words, score, raw_score) tuples, best first; words include the final 0 (</s>) of finished hypotheses."""
        _, src_mask = self.get_batch(0)
        B = src_mask.shape[1]
        max_len = max_len or 3 * src_mask.shape[0] + 2
        words = np.zeros((B, n_best, max_len), dtype=np.int64)
        lengths = np.zeros((B, n_best), dtype=np.int32)
        scores = np.zeros((B, n_best), dtype=np.float32)
        raw = np.zeros((B, n_best), dtype=np.float32)
        self.lib._ck(self.lib.c.mrn_trainer_translate(self.h, options.encode(), n_best, max_len, words.ctypes.data, lengths.ctypes.data, scores.ctypes.data, raw.ctypes.data))
        out = []
        for s in range(B):
            out.append([(words[s, r, :lengths[s, r]].tolist(), float(scores[s, r]), float(raw[s, r])) for r in range(n_best) if lengths[s, r] >= 0])
        return out

    def translate_file(self, src_path, vocab_src, vocab_trg, out_path, options=""):
        n = ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_translate_file(self.h, str(src_path).encode(), str(vocab_src).encode(), str(vocab_trg).encode(), options.encode(),
                                                           str(out_path).encode(), ctypes.byref(n)))
        return n.value

    def get_batch(self, side):
        """(indices [T, B] int64, mask [T, B] float32) of the current batch."""
        b, w = ctypes.c_int(), ctypes.c_int()
        self.lib._ck(self.lib.c.mrn_trainer_get_batch(self.h, side, None, None, 0, ctypes.byref(b), ctypes.byref(w)))
        idx = np.empty((w.value, b.value), dtype=np.int64)
        mask = np.empty((w.value, b.value), dtype=np.float32)
        self.lib._ck(self.lib.c.mrn_trainer_get_batch(self.h, side, idx.ctypes.data, mask.ctypes.data, idx.size, ctypes.byref(b), ctypes.byref(w)))
        return idx, mask

    def compute_gradients(self, keep_logits=False):
        self.lib._ck(self.lib.c.mrn_trainer_compute_gradients(self.h, int(keep_logits)))

    def update(self):
        self.lib._ck(self.lib.c.mrn_trainer_update(self.h))

    def update_shard(self):
        self.lib._ck(self.lib.c.mrn_trainer_update_shard(self.h))

    # ---- peer-memory exchange (CUDA IPC) ----
    IPC_BYTES = 3 * 64

    def ipc_export(self):
        buf = ctypes.create_string_buffer(self.IPC_BYTES)
        self.lib._ck(self.lib.c.mrn_trainer_ipc_export(self.h, buf, self.IPC_BYTES))
        return buf.raw

    def ipc_import(self, all_handles, nranks):
        assert len(all_handles) == nranks * self.IPC_BYTES
        self.lib._ck(self.lib.c.mrn_trainer_ipc_import(self.h, all_handles, nranks))

    def update_peer(self):
        self.lib._ck(self.lib.c.mrn_trainer_update_peer(self.h))

    # ---- asynchronous parameter server (options "graph-group=async") ----
    def async_init(self):
        self.lib._ck(self.lib.c.mrn_trainer_async_init(self.h))

    def async_export(self):
        buf = ctypes.create_string_buffer(64)
        self.lib._ck(self.lib.c.mrn_trainer_async_export(self.h, buf, 64))
        return buf.raw

    def async_import(self, all_handles, nranks):
        assert len(all_handles) == nranks * 64
        self.lib._ck(self.lib.c.mrn_trainer_async_import(self.h, all_handles, nranks))

    def async_update(self):
        self.lib._ck(self.lib.c.mrn_trainer_async_update(self.h))

    def async_fetch(self):
        self.lib._ck(self.lib.c.mrn_trainer_async_fetch(self.h))

    def save(self, path, with_optimizer=False):
        self.lib._ck(self.lib.c.mrn_trainer_save(self.h, str(path).encode(), int(with_optimizer)))

    def load(self, path, with_optimizer=False):
        self.lib._ck(self.lib.c.mrn_trainer_load(self.h, str(path).encode(), int(with_optimizer)))

    def cost(self):
        c = ctypes.c_float()
        self.lib._ck(self.lib.c.mrn_trainer_cost(self.h, ctypes.byref(c)))
        return c.value

    def _arena(self, fn):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self.lib._ck(fn(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def params_arena(self):
        return self._arena(self.lib.c.mrn_trainer_params)

    def grads_arena(self):
        return self._arena(self.lib.c.mrn_trainer_grads)

    def shard_grads_arena(self):
        return self._arena(self.lib.c.mrn_trainer_shard_grads)

    def arena_numpy(self, which="params"):
        ptr, n = {"params": self.params_arena, "grads": self.grads_arena, "shard_grads": self.shard_grads_arena}[which]()
        out = np.empty(n, dtype=np.float32)
        self.lib._ck(self.lib.c.mrn_memcpy_d2h(out.ctypes.data, ptr, n * 4))
        return out

    def get_tensor(self, name, grad=False):
        n = ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_get_tensor(self.h, name.encode(), int(grad), None, 0, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        self.lib._ck(self.lib.c.mrn_trainer_get_tensor(self.h, name.encode(), int(grad), out.ctypes.data, n.value, ctypes.byref(n)))
        return out

    def param_names(self):
        need = ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_param_names(self.h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        self.lib._ck(self.lib.c.mrn_trainer_param_names(self.h, buf, need.value, ctypes.byref(need)))
        out = []
        for line in buf.value.decode().strip().split("\n"):
            parts = line.split()
            out.append((parts[0], tuple(int(x) for x in parts[2:])))
        return out

    def batch_words(self):
        s, t = ctypes.c_size_t(), ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_batch_words(self.h, ctypes.byref(s), ctypes.byref(t)))
        return s.value, t.value

    def graph_kernels(self):
        k = ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_graph_kernels(self.h, ctypes.byref(k)))
        return k.value

    def stats(self):
        a, b, c, d = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        self.lib._ck(self.lib.c.mrn_trainer_stats(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return {"tape_nodes": a.value, "plans": b.value, "replays": c.value, "workspace_peak_bytes": d.value}
