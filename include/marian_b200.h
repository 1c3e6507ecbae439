/* marian_b200.h — C ABI of the B200-native hot path of Marian (v1.2.1 fork
 * tneck/marian-nmt-distributed).
 *
 * The reference has no plugin/FFI boundary: its drop-in boundary is the set of
 * C++ free functions on marian::Tensor declared in
 *   /root/reference/src/kernels/tensor_operators.h:19-397  (+ kernels/dropout.h:10-12)
 * which the graph nodes (src/graph/node_operators_*.h), optimizers
 * (src/optimizers/*.cu) and graph groups (src/training/graph_group_*.cu) call.
 * The C++ side of that boundary is kept source-compatible in
 *   marian-nmt-distributed_b200/csrc/kernels/tensor_operators.h
 * This header is the flat C view of the SAME kernels plus the training-step
 * driver, for callers that cannot link C++ (ctypes, cgo, JNI ...).  Every entry
 * cites the reference interface it stands for.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    mrn_last_error() returns the message of the calling thread's last failure.
 *    (The reference aborts the process: src/common/logging.h:43-65.)
 *  - tensors are described by mrn_tensor: a raw float pointer in DEVICE memory
 *    plus a Marian shape (rank <= 4, row-major, last dimension contiguous),
 *    exactly what marian::TensorBase carries (src/tensors/tensor.h:15-65).
 *    Shapes are right-aligned to 4-D inside the kernels like
 *    gpu::ConstantShape (src/gpu/shape.h:40-51).
 *  - all device work is asynchronous on one stream (mrn_set_stream); only the
 *    functions documented as blocking synchronise.
 *  - forward operators overwrite `out`; backward operators accumulate (+=)
 *    EXCEPT transpose_nd, shift, highway_backward, deconcatenate (they assign),
 *    as in the reference.
 */
#ifndef MARIAN_B200_H
#define MARIAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrn_tensor {
  float* data; /* device pointer */
  int rank;    /* 1..4 */
  int shape[4];
} mrn_tensor;

/* ---- library / device ------------------------------------------------- */
const char* mrn_last_error(void);
const char* mrn_backend_name(void); /* "cuda" (product) or "cpu-oracle" (test oracle build) */
int mrn_set_device(int device);
int mrn_set_stream(void* cuda_stream); /* NULL = engine-owned stream */
int mrn_synchronize(void);             /* blocking */

/* device memory helpers (cudaMalloc / cudaMemcpyAsync on the engine stream) */
int mrn_malloc(void** ptr, size_t bytes);
int mrn_free(void* ptr);
int mrn_memcpy_h2d(void* dst, const void* src, size_t bytes); /* blocking */
int mrn_memcpy_d2h(void* dst, const void* src, size_t bytes); /* blocking */
int mrn_memset_zero(void* dst, size_t bytes);

/* ---- GEMM context: replaces cublasHandle_t of Prod/ProdBatched --------- */
/* mode: 0 = fp32 SIMT (exact), 1 = packed bf16 tcgen05, 2 = bf16x3 split tcgen05 (fp32-grade, parity runs),
 *       3 = tf32 tcgen05 straight on the fp32 tensors (TMA-direct, no packing pass),
 *       4 = bf16 tcgen05 on bf16 shadow copies of the operands (written by the producing kernels /
 *           the optimizer, TMA-direct for all four transpose cases; fp32 accumulate and fp32 storage) */
int mrn_gemm_create(void** handle, int device);
int mrn_gemm_destroy(void* handle);
int mrn_gemm_set_mode(void* handle, int mode);

/* measurement hook: enable != 0 resets and starts timing every tensor-core GEMM launch (CUDA event
 * pairs on the engine stream - external event-record nodes when the step is being captured, so
 * every replay re-stamps them - or, with MRN_GEMM_SPANS=1 in the environment, in-kernel
 * %globaltimer spans); enable == 0 stops and returns total ms, algorithmic flops (2MNK) and the
 * launch count of the last eager step or graph replay. */
int mrn_gemm_profile(int enable, double* ms, double* flops, size_t* launches);
/* tuning aid: per-CTA timestamps (5 x uint64 per CTA) of the next tf32 GEMM launches; NULL disarms */
int mrn_gemm_debug_stamps(void* device_buffer);

/* Prod / ProdBatched: tensor_operators.h:295-311, .cu:543-654 */
int mrn_prod(void* gemm, mrn_tensor C, mrn_tensor A, mrn_tensor B, int transA, int transB, float beta, float scalar);
int mrn_prod_batched(void* gemm, mrn_tensor C, mrn_tensor A, mrn_tensor B, int transA, int transB, float beta, float scalar);
/* C = beta C + sum_g A_g B_g^T (n <= 3 pairs of identical shapes run as ONE K-grouped tensor-core launch in
 * tf32 mode, otherwise as the chain of accumulating products).  Replaces the chain of Prod(..., beta = 1)
 * calls the reference issues when several projections share their input (AffineNodeOp::backwardOps,
 * src/graph/node_operators_binary.h:197-214, once per q/k/v projection of Transformer::MultiHead). */
int mrn_prod_grouped_nt(void* gemm, mrn_tensor C, const mrn_tensor* As, const mrn_tensor* Bs, int n, float beta);
/* C = beta C + (A B^T) o swish'(H), swish'(h) = s(h) (1 + h (1 - s(h))): the input gradient of an affine layer
 * applied to swish(H), delivered straight into the adjoint of H.  Replaces Prod(..., false, true, 1.0) of
 * AffineNodeOp::backwardOps followed by SwishNodeOp::backwardOps (src/graph/node_operators_unary.h, the
 * "swish" functor of src/functional/predicates.h).  tf32 mode only; returns an error otherwise. */
int mrn_prod_swish_grad_nt(void* gemm, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor H, float beta);
/* The two products above with the bias gradients they deliver in the training step: col_sums[g] (+)= column sums of
 * A_g (A = the adjoint of an affine node, so its column sums are the gradient of that node's bias: Add(_1, bias->grad(),
 * adj) of AffineNodeOp::backwardOps, src/graph/node_operators_binary.h:208-212, reference kernel gAddGeneric), taken
 * from the A tiles while they stream through the tensor-core kernel.  Tensor-core modes; elsewhere the sums are
 * produced by the column-sum kernel. */
int mrn_prod_grouped_nt_sums(void* gemm, mrn_tensor C, const mrn_tensor* As, const mrn_tensor* Bs, int n, float beta, const mrn_tensor* col_sums);
int mrn_prod_swish_grad_nt_sums(void* gemm, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor H, float beta, mrn_tensor col_sum);
/* C_i = beta C_i + op(A) B_i (+ bias_i), i < n <= 3: products that share their A operand - the query / key / value
 * projections of Transformer::MultiHead (src/models/transformer.h:194-261: three affine() calls on the same input) and
 * their weight gradients (AffineNodeOp::backwardOps, three Prod(.., true, false, 1.0) with the same x) - as ONE launch in
 * the bf16 shadow mode (*fused = 1), otherwise as the n single products the reference issues (*fused = 0).
 * biases may be NULL. */
int mrn_prod_shared_a(void* gemm, const mrn_tensor* Cs, mrn_tensor A, const mrn_tensor* Bs, const mrn_tensor* biases, int n, int transA, float beta, int* fused);
/* AffineNodeOp forward (Prod + Add(_1, val, bias)): node_operators_binary.h:172-186 */
int mrn_prod_affine(void* gemm, mrn_tensor C, mrn_tensor A, mrn_tensor B, mrn_tensor bias);

/* ---- element-wise family: Element / Add / Reduce templates ------------- */
/* tensor_operators.h:26-274.  The functor is selected by name from the set the
 * graph nodes use (node_operators_unary.h / _binary.h), e.g. "plus" (_1 = _2 + _3),
 * "mult", "minus", "div", "swish", "tanh3", "logit", "relu", "scale" (_1 = c * _2),
 * "shift" (_1 = _2 + c), "neg", "exp", "log", "sqrt", "square".
 * n_in inputs (<= 3), `c` is the captured scalar where the functor has one. */
int mrn_element(const char* functor, mrn_tensor out, const mrn_tensor* ins, int n_in, float c);
/* out += scale * reduce_or_broadcast(f(ins...)); functors: "id" (_1), "mult" (_1*_2),
 * "neg", "tanh_grad" (_1*(1-_2*_2)), "swish_grad", "div_grad_b", "scale" (c*_1) ... */
int mrn_add(const char* functor, float scale, mrn_tensor out, const mrn_tensor* ins, int n_in, float c);

/* ---- row operators ------------------------------------------------------ */
/* Softmax / LogSoftmax (+Grad): tensor_operators.h:278-282, .cu:202-519. mask may be NULL */
int mrn_softmax(mrn_tensor out, mrn_tensor in, const mrn_tensor* mask);
int mrn_logsoftmax(mrn_tensor out, mrn_tensor in);
int mrn_softmax_grad(mrn_tensor grad, mrn_tensor adj, mrn_tensor val);
int mrn_logsoftmax_grad(mrn_tensor grad, mrn_tensor adj, mrn_tensor val);
/* CrossEntropyPick(+Backward): tensor_operators.h:290-291, .cu:1115-1283; pick holds labels as floats */
int mrn_cross_entropy_pick(mrn_tensor out, mrn_tensor in, mrn_tensor pick);
int mrn_cross_entropy_pick_backward(mrn_tensor out, mrn_tensor adj, mrn_tensor in, mrn_tensor pick);
/* LayerNormalization(+Grad): tensor_operators.h:351-364, .cu:1447-1674; beta / grad_beta may be NULL */
int mrn_layer_norm(mrn_tensor out, mrn_tensor in, mrn_tensor gamma, const mrn_tensor* beta, float eps);
int mrn_layer_norm_grad(mrn_tensor grad_x, mrn_tensor grad_gamma, const mrn_tensor* grad_beta, mrn_tensor adj, mrn_tensor y, mrn_tensor x, mrn_tensor gamma, const mrn_tensor* beta, float eps);
/* layer_norm(in + residual): the "add residual, then normalise" tail of a Transformer sub-layer
 * (src/models/transformer.h:111-123, PlusNodeOp + LayerNormalizationOp in the reference) as one operator;
 * grad_x and grad_residual both receive d(in + residual) (accumulating). */
int mrn_residual_layer_norm(mrn_tensor out, mrn_tensor in, mrn_tensor residual, mrn_tensor gamma, mrn_tensor beta, float eps);
int mrn_residual_layer_norm_grad(mrn_tensor grad_x, mrn_tensor grad_residual, mrn_tensor grad_gamma, mrn_tensor grad_beta, mrn_tensor adj, mrn_tensor y, mrn_tensor x, mrn_tensor residual, mrn_tensor gamma, mrn_tensor beta, float eps);
/* Fused multi-head attention core = the node sequence of Transformer::MultiHead / Attention
 * (src/models/transformer.h:58-77,153-261: SplitHeads, bdot, + mask, softmax, bdot, JoinHeads).
 * q [B,Tq,d], k/v [B,Tk,d], additive mask with B*Tk or B*Tq*Tk elements (may be NULL),
 * probs [B,heads,Tq,Tk] (softmax output, kept for the backward pass).  The gradients ACCUMULATE.
 * exact != 0: 3xTF32 products (fp32-grade); exact == 0: plain tf32 operands. */
int mrn_multi_head_attention(mrn_tensor out, mrn_tensor probs, mrn_tensor q, mrn_tensor k, mrn_tensor v, const mrn_tensor* mask, int heads, float scale, int exact);
int mrn_multi_head_attention_grad(mrn_tensor dq, mrn_tensor dk, mrn_tensor dv, mrn_tensor adj, mrn_tensor out, mrn_tensor probs, mrn_tensor q, mrn_tensor k, mrn_tensor v, int heads, float scale, int exact);
/* Att / AttBack: tensor_operators.h:342-349, .cu:1307-1445 */
int mrn_att(mrn_tensor out, mrn_tensor va, mrn_tensor context, mrn_tensor state);
int mrn_att_back(mrn_tensor g_va, mrn_tensor g_context, mrn_tensor g_state, mrn_tensor va, mrn_tensor context, mrn_tensor state, mrn_tensor adj);

/* ---- recurrent cells ---------------------------------------------------- */
/* GRUFastForward/Backward: tensor_operators.h:335-340, .cu:934-1113.
 * inputs = {state, xW, sU, b[, mask]} ; outputs (grads) may have data == NULL */
int mrn_gru_fast_forward(mrn_tensor out, const mrn_tensor* inputs, int n_inputs, int final);
int mrn_gru_fast_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n_inputs, mrn_tensor adj, int final);
/* LSTMCell/Output Forward/Backward: tensor_operators.h:326-333, .cu:1749-2031 */
int mrn_lstm_cell_forward(mrn_tensor out, const mrn_tensor* inputs, int n_inputs);
int mrn_lstm_output_forward(mrn_tensor out, const mrn_tensor* inputs, int n_inputs);
int mrn_lstm_cell_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n_inputs, mrn_tensor adj);
int mrn_lstm_output_backward(const mrn_tensor* outputs, const mrn_tensor* inputs, int n_inputs, mrn_tensor adj);
/* HighwayForward/Backward: tensor_operators.h:372-383, .cu:2033-2104 */
int mrn_highway_forward(mrn_tensor out, mrn_tensor in1, mrn_tensor in2, mrn_tensor t);
int mrn_highway_backward(mrn_tensor out1, mrn_tensor out2, mrn_tensor outt, mrn_tensor in1, mrn_tensor in2, mrn_tensor t, mrn_tensor adj);

/* ---- data movement ------------------------------------------------------ */
/* TransposeND: tensor_operators.h:72, .cu:164-200; axes has in.rank entries */
int mrn_transpose_nd(mrn_tensor out, mrn_tensor in, const int* axes);
/* Concatenate / Deconcatenate: tensor_operators.h:86-88, .cu:35-162 */
int mrn_concatenate(mrn_tensor out, const mrn_tensor* ins, int n, int axis);
int mrn_deconcatenate(const mrn_tensor* outs, int n, mrn_tensor in, int axis);
/* CopyRows / PasteRows: tensor_operators.h:318-320, .cu:656-746; indices: int32 in DEVICE memory */
int mrn_copy_rows(mrn_tensor out, mrn_tensor in, const int* device_indices, size_t n);
int mrn_paste_rows(mrn_tensor out, mrn_tensor in, const int* device_indices, size_t n);
/* Shift: tensor_operators.h:366, .cu:1676-1707; shift has in.rank entries */
int mrn_shift(mrn_tensor out, mrn_tensor in, const int* shift, int invert);

/* ---- norms and optimizer ------------------------------------------------ */
/* L2Norm: tensor_operators.h:276, .cu:1286-1305 (blocking: returns the value) */
int mrn_l2norm(mrn_tensor in, float* result);
/* Norm::clip + Adam::updateImpl fused: optimizers/clippers.cu:12-17, optimizers.cu:43-73.
 * t = 1-based step; grad_scale multiplies every gradient (1/N of a summed shard);
 * clip_norm <= 0 disables clipping. */
int mrn_adam_step(mrn_tensor params, mrn_tensor grads, mrn_tensor mt, mrn_tensor vt, float eta, float beta1, float beta2, float eps, int t, float grad_scale, float clip_norm);
/* Norm::clip + Sgd::updateImpl / Adagrad::updateImpl fused: optimizers/optimizers.cu:7-41
 * (p -= eta g;   gt += g^2, p -= eta / (sqrt(gt) + eps) g), same grad_scale / clip_norm meaning. */
int mrn_sgd_step(mrn_tensor params, mrn_tensor grads, float eta, float grad_scale, float clip_norm);
int mrn_adagrad_step(mrn_tensor params, mrn_tensor grads, mrn_tensor gt, float eta, float eps, float grad_scale, float clip_norm);
/* Dropout mask: kernels/dropout.cu:25-42 (Bernoulli(1 - p) keep mask scaled by 1 / (1 - p)); the VALUES
 * come from this library's counter-based generator (the reference pins no cuRAND values). */
int mrn_dropout(mrn_tensor mask, float drop_prob, unsigned long long seed);

/* ======================================================================== */
/* Training-step driver: ExpressionGraph + model + GraphGroup behind a handle
 * (reference: Train<GraphGroup>::run's hot loop, src/training/training.h:51-57;
 * SingletonGraph::execute graph_group_singleton.cu:21-66; SyncGraphGroup::execute
 * graph_group_sync.cu:42-188).
 * options: "key=value;key=value" over the reference's defaults
 * (src/common/config_parser.cpp:214-467), e.g.
 *   "type=transformer;dim-vocabs=32000,32000;enc-depth=6;dec-depth=6;gemm-mode=1"
 * rank/nranks > 1 selects the sharded SyncGraphGroup of this rank. */
int mrn_trainer_create(void** trainer, const char* options, int device, int rank, int nranks);
int mrn_trainer_destroy(void* trainer);

/* Batch in the reference's CorpusBatch layout (src/data/corpus.h:49-205):
 * per side time-major indices[t*B+b] (int64) and mask[t*B+b] in {0,1}; HOST memory. */
int mrn_trainer_set_batch(void* trainer, int batch_size, int src_len, const int64_t* src_idx, const float* src_mask, int trg_len, const int64_t* trg_idx, const float* trg_mask);
/* Synthetic bitext of SURVEY.md 8d generated inside the library (same RNG stream on
 * both builds): padded = 0 dense / 1 padded+sorted.  Advances the corpus. */
int mrn_trainer_next_synthetic_batch(void* trainer, int batch_size, int max_len_src, int max_len_trg, int padded, int split_rank, int split_n);

/* Text corpus in front of the hot path (reference: Vocab src/data/vocab.cpp, Corpus src/data/corpus.cpp:30-230,
 * BatchGenerator src/data/batch_generator.h:39-160): one text file per side, vocabularies as YAML maps word -> id
 * ("</s>" = 0, "<unk>" = 1; NULL or "" = <corpus>.yml, created from the corpus by falling frequency if missing),
 * options "mini-batch=64;maxi-batch=100;maxi-batch-sort=trg|src|none;mini-batch-words=0;max-length=50;
 * max-length-crop=false;right-left=false;shuffle=true;seed=1234".  A host thread reads, sorts and assembles the
 * mini-batches two ahead of the device.  next_corpus_batch makes the next mini-batch the current batch (this rank's
 * split when nranks > 1); *has_batch = 0 marks the end of an epoch. */
int mrn_trainer_open_corpus(void* trainer, const char* src_path, const char* trg_path, const char* vocab_src, const char* vocab_trg, const char* options);
int mrn_trainer_next_corpus_batch(void* trainer, int* has_batch);
/* Cross-entropy validation (reference CrossEntropyValidator, src/training/validator.h:108-176): forward passes of an
 * inference-mode model (no dropout, cost-type ce-sum) over a held-out text corpus on the trainer's own parameters;
 * *metric follows "cost-type" of the options (cross-entropy / ce-mean: cost per sentence, ce-mean-words, perplexity,
 * ce-sum); cost_sum / sentences / target_words receive the totals (may be NULL).  options as for
 * mrn_trainer_open_corpus plus "valid-mini-batch", "valid-max-length".  Blocking; parameters are not changed. */
int mrn_trainer_validate(void* trainer, const char* src_path, const char* trg_path, const char* vocab_src, const char* vocab_trg, const char* options, float* metric, float* cost_sum,
                         size_t* sentences, size_t* target_words);
/* Beam-search decoding on the trainer's parameters (inference mode: no dropout).
 * Replaces BeamSearch::search + History::NBest (src/translator/beam_search.h:91-225, src/translator/history.h:36-66)
 * with NthElement::getNBestList (src/translator/nth_element.cu:270-402) underneath.
 * options: "beam-size=12;normalize=0;allow-unk=false;beam-fused-nth=true" (the reference's option names; beam-fused-nth
 * selects the fused logsoftmax + n-best kernel instead of the reference's node sequence - same result).
 * mrn_trainer_translate: SOURCE side of the current batch; for sentence s and rank r < n_best, slot = s*n_best + r:
 * lengths[slot] words (incl. the final </s> = 0; -1 if there are fewer than n_best hypotheses) at words[slot*max_len ...],
 * scores[slot] = log-probability / length^normalize, raw_scores[slot] (may be NULL) = log-probability. */
int mrn_trainer_translate(void* trainer, const char* options, int n_best, int max_len, int64_t* words, int* lengths, float* scores, float* raw_scores);
/* Text file -> translations, one line per input line in corpus order ("n-best=true": "id ||| words ||| F0= cost ||| cost"
 * lines), as Translate<BeamSearch>::run does (src/translator/translator.h:22-108, output_collector.cpp).  Further
 * options: "mini-batch=1;maxi-batch=1;max-length=1000". */
int mrn_trainer_translate_file(void* trainer, const char* src_path, const char* vocab_src, const char* vocab_trg, const char* options, const char* out_path, size_t* sentences);
/* n best (value, flat index) pairs of every range [range_first[i], range_first[i+1]) of a device array, best first
 * (NthElement::getNBestList, src/translator/nth_element.cu:343-361); n_i = cum_n[i+1]-cum_n[i].  Blocking. */
int mrn_nth_element_ranges(mrn_tensor scores, const int* range_first, const int* cum_n, int ranges, float* out_costs, unsigned* out_keys);
/* The same selection fused with log-softmax, previous-cost add and per-sentence regrouping, from raw logits
 * [beam, 1, batch, V] (kernels/nth_element.cu): replaces the node sequence of src/translator/beam_search.h:163-196. */
int mrn_nth_element_logsoftmax(mrn_tensor logits, const float* prev_costs, int dim_batch, int beam, int n, int first, int suppress_word, float* out_costs, unsigned* out_keys);

/* the current batch as host arrays in the SubBatch layout (time-major [T, B]); side 0 = source, 1 = target;
 * indices / mask may be NULL to query batch_size and width */
int mrn_trainer_get_batch(void* trainer, int side, int64_t* indices, float* mask, size_t capacity, int* batch_size, int* width);

/* forward + backward of the current batch (CUDA-graph replay after the first
 * occurrences of a shape).  Asynchronous. keep_logits != 0 keeps the logits node. */
int mrn_trainer_compute_gradients(void* trainer, int keep_logits);
/* single-process update: clip + optimizer over the whole arena. Asynchronous. */
int mrn_trainer_update(void* trainer);
/* sharded update of this rank's shard from the summed gradient shard
 * (mrn_trainer_shard_grads_ptr), then invalidates packed weights. Asynchronous. */
int mrn_trainer_update_shard(void* trainer);
/* blocking: waits for the stream, returns the cost of the last batch */
/* Peer-memory exchange (replaces the reference's gather / add / update / scatter loop of
 * SyncGraphGroup::execute, src/training/graph_group_sync.cu:125-151, for ranks on one node):
 * every rank exports CUDA IPC handles of {parameter arena, gradient arena, signal pad}
 * (3 x 64 bytes), the host exchanges them (any transport), every rank imports all of them
 * (nranks x 3 x 64 bytes, rank-major), then mrn_trainer_update_peer() runs
 * {barrier, gather-reduce by peer loads, clip + Adam with peer stores, barrier} on the engine
 * stream after mrn_trainer_compute_gradients().  Adam only. */
int mrn_trainer_ipc_export(void* trainer, unsigned char* handles, size_t capacity);
int mrn_trainer_ipc_import(void* trainer, const unsigned char* all_handles, int nranks);
int mrn_trainer_update_peer(void* trainer);
/* Asynchronous SGD with a sharded parameter server = the reference's AsyncGraphGroup
 * (src/training/graph_group_async.cu:16-250; fetchParams / pushGradients under per-shard locks) for
 * one process per GPU: create the trainer with "graph-group=async[;optimizer-delay=tau]", set a batch,
 * mrn_trainer_async_init() (builds the parameters, allocates this rank's master shard), exchange the
 * 64-byte IPC handles (export / import, rank-major), then every mrn_trainer_async_update() is
 * {fetch every tau steps, forward+backward, push every tau steps}.  nranks = 1 works without peers
 * after importing the own handle.  Adam only. */
int mrn_trainer_async_init(void* trainer);
int mrn_trainer_async_export(void* trainer, unsigned char* handle, size_t capacity);
int mrn_trainer_async_import(void* trainer, const unsigned char* all_handles, int nranks);
int mrn_trainer_async_update(void* trainer);
int mrn_trainer_async_fetch(void* trainer);
int mrn_trainer_cost(void* trainer, float* cost);

/* Checkpoint / resume in the reference's wire format (ExpressionGraph::save / load,
 * src/graph/expression_graph.h:442-502; EncoderDecoder::save, src/models/encdec.h:201-229): one .npz with
 * every parameter as float32 under its Marian name + the char array "special:model.yml".  load() goes into a
 * FRESH trainer before its first step (the parameters are created from the file, the model code finds them by
 * name) and checks the stored model description against the trainer's options.  with_optimizer != 0 also
 * writes / reads "<path>.optimizer.npz" (Adam moments per parameter name + step counter; the reference has no
 * optimizer checkpoint).  Blocking. */
int mrn_trainer_save(void* trainer, const char* path, int with_optimizer);
int mrn_trainer_load(void* trainer, const char* path, int with_optimizer);

/* flat arenas (src/graph/parameters.h:58-80): device pointers + element counts */
int mrn_trainer_params(void* trainer, float** ptr, size_t* elements);
int mrn_trainer_grads(void* trainer, float** ptr, size_t* elements);
int mrn_trainer_shard_grads(void* trainer, float** ptr, size_t* elements);
/* copies of named tensors to HOST memory (blocking); name = parameter name,
 * or "logits" / "cost"; returns the element count in *elements (buffer may be NULL to query) */
int mrn_trainer_get_tensor(void* trainer, const char* name, int want_grad, float* host_buffer, size_t capacity, size_t* elements);
/* newline separated "name rank d0 d1 .." list of parameters in arena order */
int mrn_trainer_param_names(void* trainer, char* buffer, size_t capacity, size_t* needed);
/* words in the current batch: source-side (reference's log line) and source+target */
int mrn_trainer_batch_words(void* trainer, size_t* src_words, size_t* total_words);
/* statistics: number of tape nodes of the last eager build, captured plans, replays */
int mrn_trainer_stats(void* trainer, size_t* tape_nodes, size_t* plans, size_t* replays, size_t* workspace_bytes);

/* kernel nodes in the most recently captured step graph (0 before any capture) */
int mrn_trainer_graph_kernels(void* trainer, size_t* kernels);

/* Runs one of the reference's unit-test graphs (src/tests/*.cpp) through the
 * graph API and returns the values the reference test asserts on.  Used by the
 * golden-vector tests; see tests/cpp/graph_golden.cpp for the case names. */
int mrn_test_golden(const char* test_case, float* out, size_t capacity, size_t* count);

#ifdef __cplusplus
}
#endif
#endif /* MARIAN_B200_H */
