// THE DROP-IN BOUNDARY of the hot path: the tensor operators every graph node,
// optimizer and graph group is written against.
//
// Names, argument order and semantics follow the reference's
// src/kernels/tensor_operators.h:19-397 (+ src/kernels/dropout.h:10-12):
// forward operators overwrite `out`, backward operators accumulate (+=) into
// their gradient outputs EXCEPT TransposeND, Shift, HighwayBackward and
// Deconcatenate, which assign.  Implementations:
//   * product build: hand-written sm_100a CUDA in csrc/kernels/*.cu, all
//     asynchronous on device::currentStream();
//   * test oracle:   oracle/cpu/tensor_operators_cpu.cpp (CPU restatement).
// The element-wise templates Element/Add/Reduce live in kernels/element.h
// (the oracle shadows that header with host loops).
//
// The flat C ABI over the same kernels is include/marian_b200.h.
#pragma once

#include <vector>

#include "common/definitions.h"
#include "common/shape.h"
#include "functional/functional.h"
#include "tensors/allocator.h"
#include "tensors/tensor.h"

namespace marian {

// GEMM context: replaces the cublasHandle_t argument of the reference's
// Prod/ProdBatched (tensor_operators.h:295-311).  Owns the arithmetic mode and
// the scratch for bf16-packed operands.
enum class GemmMode : int {
  FP32 = 0,    // fp32 SIMT contraction (exact-mode, used for 1e-4 parity runs)
  BF16 = 1,    // tcgen05 bf16 operands, fp32 accumulate in TMEM (throughput mode)
  BF16X3 = 2,  // tcgen05 with hi/lo bf16 split operands (3 products), ~fp32 accuracy
  TF32 = 3,    // tcgen05 kind::tf32 directly on the fp32 tensors (no packing pass; throughput mode)
  BF16S = 4,   // tcgen05 kind::f16 on bf16 SHADOW copies of the operands, TMA-direct in all four
               // transpose cases (kernels/shadow.h); fp32 storage and accumulation (headline mode)
};
struct GemmContext;
typedef GemmContext* GemmHandle;
GemmHandle createGemmContext(int deviceId);
void destroyGemmContext(GemmHandle);
void setGemmMode(GemmHandle, GemmMode);
GemmMode getGemmMode(GemmHandle);
// Packed-operand cache: bf16 copies of GEMM operands are keyed by (pointer,
// layout) and reused until invalidated; the graph invalidates at the start of
// forward and after the optimizer step.
void gemmInvalidateCache(GemmHandle);
// Sources inside [ptr, ptr+bytes) (the parameter arena) are packed once per
// step: their bf16 copies survive until the next gemmInvalidateCache().
void gemmSetStableRange(GemmHandle, const void* ptr, size_t bytes);
// BF16S mode keeps a bf16 copy of the whole parameter arena (the stable range).
//   gemmParamShadowFor(h, t): where an optimizer kernel should store the bf16 copy of parameter
//       tensor / shard `t` (nullptr when the mode is off or t is not inside the arena);
//   gemmParamsUpdated(h, shadowWritten): the parameters changed; shadowWritten = the update wrote
//       the bf16 copy of the WHOLE arena itself;
//   gemmPrepareStep(h): before a forward pass / graph replay: refreshes the bf16 arena copy by one
//       flat conversion pass if it is stale.  All three are no-ops in the other modes and on the oracle.
// shadow-only tensors (kernels/shadow.h): allow = false keeps every fp32 tensor complete for this step (a step whose
// logits / adjoints the host wants to read back)
void gemmAllowShadowOnly(GemmHandle, bool allow);
void* gemmParamShadowFor(GemmHandle, const Tensor& t);
void gemmParamsUpdated(GemmHandle, bool shadowWritten);
void gemmPrepareStep(GemmHandle);

// Tuning aid (scripts/gemm_stamps.py): per-CTA %globaltimer stamps of the tf32 kernel.
void gemmDebugStamps(unsigned long long* deviceBuffer);
// Per-launch CUDA-event timing of the tensor-core GEMM (eager steps only).
void gemmProfile(int enable, double* ms, double* flops, size_t* launches);

bool IsNan(Tensor in);

void TransposeND(Tensor out, Tensor in, const std::vector<int>& vAxis);

void Select(Ptr<Allocator> allocator, Tensor out, Tensor in, int axis, const std::vector<size_t>&);
void Insert(Ptr<Allocator> allocator, Tensor out, Tensor in, int axis, const std::vector<size_t>&);

void Concatenate(Tensor out, const std::vector<Tensor>& inputs, int ax);
void Deconcatenate(std::vector<Tensor>& outputs, const Tensor in, int ax);

float L2Norm(Tensor in);
// Asynchronous variant: writes sum(x^2) (NOT the root) to a device scalar.
void SumSquares(Tensor outScalar, Tensor in);

void Softmax(Tensor out, Tensor in, Tensor mask = nullptr);
void LogSoftmax(Tensor out, Tensor in);
void SoftmaxGrad(Tensor grad, Tensor adj, Tensor val);
void LogSoftmaxGrad(Tensor grad, Tensor adj, Tensor val);

// `stats` (optional, [rows, 2]): the forward pass stores each row's max and sum of exponentials,
// the backward pass then reads the logits once instead of twice (410 MB per read at config B).
void CrossEntropyPick(Tensor out, Tensor in, Tensor pick, Tensor stats = nullptr);
void CrossEntropyPickBackward(Tensor out, Tensor adj, Tensor a, Tensor pick, Tensor stats = nullptr);

void Prod(GemmHandle handle, Tensor C, const Tensor A, const Tensor B, bool transA, bool transB, float beta = 0, float scalar = 1);
// C = beta C + sum_g A_g B_g^T: the input gradient of several projections of one tensor, as ONE K-grouped launch
// colSums (optional, one per pair): colSums[g] += column sums of A_g - the bias gradients, summed from the A tiles the product streams anyway
void ProdGroupedNT(GemmHandle handle, Tensor C, const std::vector<Tensor>& As, const std::vector<Tensor>& Bs, float beta = 0, const std::vector<Tensor>& colSums = {});
// C_g = beta C_g + op(A) B_g (+ bias_g): 2 or 3 products that share their A operand as ONE launch (the q / k / v projections of
// an attention block, their weight gradients).  false = not applicable here, nothing was done (issue them one by one).
bool ProdSharedA(GemmHandle handle, const std::vector<Tensor>& Cs, const Tensor A, const std::vector<Tensor>& Bs, const std::vector<Tensor>& biases, bool transA, float beta);
bool ProdColumnSumsFusable(GemmHandle handle, const Tensor A);
// Column sums (bias gradients) that a product did not take from its own A tiles are queued by the product and issued
// here, on the side stream, as passes over the bf16 copies of those operands (kernels/gemm.cu: gColumnSumsBf16).  The
// graph calls this from the bias-gradient closure of the node - i.e. BEHIND the node's weight-gradient product on the
// side stream - and once more at the end of the backward sweep.
void ProdFlushColumnSums(GemmHandle handle);
// C = beta C + (A B^T) o swish'(H): "affine after swish" backward in the product's epilogue (tf32 tensor-core path only)
bool ProdSwishGradFusable(GemmHandle handle, const Tensor C, const Tensor A, const Tensor B, const Tensor H);
void ProdSwishGradNT(GemmHandle handle, Tensor C, const Tensor A, const Tensor B, const Tensor H, float beta = 0, Tensor colSum = nullptr);
void ProdBatched(GemmHandle handle, Tensor C, const Tensor A, const Tensor B, bool transA, bool transB, float beta = 0, float scalar = 1);
// Affine = Prod + bias row broadcast in the GEMM epilogue (the reference's
// AffineNodeOp issues Prod then Add(_1, val, bias): node_operators_binary.h:172-186).
void ProdAffine(GemmHandle handle, Tensor C, const Tensor A, const Tensor B, const Tensor bias);

// The reference takes std::vector<size_t> and cudaMallocs a device copy per
// call (tensor_operators.cu:679-746).  Here indices are int32 in device memory
// (uploaded once per graph through pinned staging); the vector overloads
// upload and forward.
void CopyRows(Tensor out, const Tensor in, const int* deviceIndices, size_t n);
void PasteRows(Tensor out, const Tensor in, const int* deviceIndices, size_t n);
void CopyRows(Tensor out, const Tensor in, const std::vector<size_t>& indices);
void PasteRows(Tensor out, const Tensor in, const std::vector<size_t>& indices);
void CopyCols(Tensor out, const Tensor in, const std::vector<size_t>& indices);
void PasteCols(Tensor out, const Tensor in, const std::vector<size_t>& indices);

void LSTMCellForward(Tensor out, std::vector<Tensor> inputs);
void LSTMOutputForward(Tensor out, std::vector<Tensor> inputs);
void LSTMCellBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj);
void LSTMOutputBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj);

void GRUFastForward(Tensor out, std::vector<Tensor> inputs, bool final = false);
void GRUFastBackward(std::vector<Tensor> outputs, std::vector<Tensor> inputs, Tensor adj, bool final = false);

void Att(Tensor out, Tensor va, Tensor context, Tensor state);
void AttBack(Tensor gva, Tensor gContext, Tensor gState, Tensor va, Tensor context, Tensor state, Tensor adj);

// Fused multi-head scaled-dot-product attention on [B, T, heads*dk] projections:
//   out = JoinHeads(softmax(scale * SplitHeads(q) SplitHeads(k)^T + mask) SplitHeads(v))
// i.e. the reference's Transformer::MultiHead core (src/models/transformer.h:153-192 with the
// SplitHeads/JoinHeads of :58-77) as ONE operator.  `mask` is ADDITIVE (0 / -99999999) and holds
// B*Tk (key mask) or B*Tq*Tk elements; `probs` [B, heads, Tq, Tk] receives the softmax output
// for the backward pass (may be null in inference).  Grad ACCUMULATES into dq/dk/dv unless the
// respective tensor is lazily zero (then it assigns).
bool AttentionFusable(int Tq, int Tk, int dimModel, int heads);
// `exact`: the products inside run as 3xTF32 (fp32-grade) when true, as plain tf32 when false
// (the graph node passes false in the tf32 / bf16 GEMM modes).
void MultiHeadAttention(Tensor out, Tensor probs, const Tensor q, const Tensor k, const Tensor v, const Tensor mask, int heads, float scale, bool exact = true);
void MultiHeadAttentionGrad(Tensor dq, Tensor dk, Tensor dv, const Tensor adj, const Tensor out, const Tensor probs, const Tensor q, const Tensor k, const Tensor v, int heads, float scale, bool exact = true);

// `residual` (optional, this repo's addition): normalise in + residual, i.e. the "add the
// residual, then layer-norm" tail of every Transformer sub-layer (src/models/transformer.h:97-126)
// in one pass; the gradient then also flows into gradResidual (accumulating unless lazily zero).
bool LayerNormResidualFusable(int cols);
void LayerNormalization(Tensor out, Tensor in, Tensor gamma, Tensor beta, float eps = 1e-9, Tensor residual = nullptr);
void LayerNormalizationGrad(Tensor gradX, Tensor gradGamma, Tensor gradBeta, Tensor adj, Tensor y, Tensor x, Tensor gamma, Tensor beta, float eps = 1e-9, Tensor residual = nullptr, Tensor gradResidual = nullptr);

void Shift(Tensor out, Tensor in, Shape shift, bool invert = false);

void HighwayForward(Tensor out, const Tensor in1, const Tensor in2, const Tensor t);
void HighwayBackward(Tensor out1, Tensor out2, Tensor outt, const Tensor in1, const Tensor in2, const Tensor t, const Tensor adj);

// Inverted-dropout mask: Bernoulli(1-p)/(1-p) (reference: kernels/dropout.cu:34-42,
// cuRAND XORWOW there; here a counter-based Philox-style hash keyed by `seed`).
// `epoch` (optional): device counter mixed into the seed.  A step that is captured once and
// replayed as a CUDA graph has its seeds baked in; the graph bumps *epoch once per forward pass
// (DropoutEpochBump, also captured), so every replay draws fresh masks.
void Dropout(Tensor mask, float dropProb, uint64_t seed, const uint64_t* epoch = nullptr);
void DropoutEpochBump(uint64_t* epoch);

// Fused optimizer kernels over flat parameter tensors (reference:
// src/optimizers/optimizers.cu:7-73 issues 1-3 Element passes per update and
// clippers.cu:12-17 a separate norm + scale pass).
// gradScale is applied to every gradient element as it is read (clipping and/or
// 1/N averaging); when `normSq` is given the clip factor is computed ON DEVICE
// from *normSq: scale = gradScale * min(1, clipNorm / sqrt(*normSq * gradScale^2)).
struct AdamArgs {
  float eta, beta1, beta2, eps;
  float denom1, denom2;  // 1 - beta1^t, 1 - beta2^t
  float gradScale;       // multiplies every gradient before use
  float clipNorm;        // <= 0: no clipping
  void* shadow{nullptr};  // != null: bf16 copy of the updated parameters is stored here as well (BF16S mode)
};
// Peer-memory exchange (kernels/exchange.cu).  A PeerTable holds, per rank of the node, the
// address at which THIS process sees that rank's buffer (own buffer: the local pointer).
struct PeerTable {
  void* ptr[8];
};
// Where an optimizer kernel additionally stores the updated parameters: element i of the shard
// goes to ptr[r] + offset + i for every rank r != self (all-gather by peer stores).
struct PeerStores {
  PeerTable params;
  int nranks{0};
  int self{0};
  size_t offset{0};
};
// Piece-wise exchange (overlapped with the backward sweep, training/graph_group.h): every rank owns piece `rank` of
// EVERY reference shard (a reference shard = one contiguous 1/N of the arena, the unit the reference clips by).
// A PieceList names the pieces a rank handles in one phase: piece k covers arena elements [off[k], off[k] + len)
// of reference shard shard[k]; its summed gradient and Adam state live at element state[k] of the rank-local
// scratch / moment tensors.
struct PieceList {
  size_t off[8];
  size_t state[8];
  int shard[8];
  int count{0};
  size_t len{0};
};
// Signal pad layout (4 KB per rank, mapped into all ranks): ints [0, 8) barrier epochs; floats at byte 1024:
// partial sums of squares [phase 2][source rank 8][reference shard 8].
constexpr size_t kSignalPadBytes = 4096;
// sums[state[k] + i] = sum_r grads_r[off[k] + i]; partialSq[shard[k]] += sum of squares of piece k (local device floats)
// background: the phase runs next to the backward sweep (small grids) instead of alone on the device
void PeerGatherReducePieces(Tensor sums, float* partialSq, const PeerTable& grads, int nranks, const PieceList& pieces, bool background);
// every rank's partialSq[0..nranks) -> slot [phase][rank][*] of every rank's signal pad (peer stores)
void PeerPublishPartials(const float* partialSq, const PeerTable& pads, int rank, int nranks, int phase);
// clip by the norm of the whole reference shard (sum over source ranks of the published partials in the OWN pad), 1/N,
// Adam on the pieces; new parameters to the local arena and by peer stores into every replica
void AdamUpdatePieces(const PeerTable& params, void* ownPad, int rank, int nranks, int phase, Tensor sums, Tensor mt, Tensor vt, const AdamArgs& args, const PieceList& pieces, bool background);
void PeerBarrier(const PeerTable& pads, int rank, int nranks, int epoch);
// tuning aid: one thread writes %globaltimer (ns) to *slot on the current stream
void DeviceTimeStamp(unsigned long long* slot);
// shardSum[i] = sum_r grads_r[offset + i];  normSq = sum_i shardSum[i]^2 (same pass)
void PeerGatherReduce(Tensor shardSum, Tensor normSq, const PeerTable& grads, int nranks, size_t offset);

// Asynchronous parameter server (training/graph_group.h AsyncGraphGroup, kernels/exchange.cu).
// A master block = [int lock, int steps, pad to 256 B][p | m | v] with `shard` floats each, on
// its owner GPU, mapped into every rank.  ShardLock spins on the lock with system-scope atomics;
// countStep also increments the shard's Adam step counter and leaves it in *stepsOut (local).
void ShardLock(void* masterBlock, bool countStep, int* stepsOut);
void ShardUnlock(void* masterBlock);
// clip (by *normSq, the squared norm of the local slice) + Adam on the remote master shard;
// args.denom1/denom2 are ignored: the bias correction uses the shard's own step counter *steps.
void AdamUpdateRemote(void* masterBlock, size_t shardElements, const float* gradSlice, const AdamArgs& args, const int* steps, Tensor normSq);

void AdamUpdate(Tensor params, Tensor grads, Tensor mt, Tensor vt, const AdamArgs& args, Tensor normSq = nullptr, const PeerStores* peers = nullptr);
void SgdUpdate(Tensor params, Tensor grads, float eta, float gradScale, float clipNorm, Tensor normSq = nullptr);
void AdagradUpdate(Tensor params, Tensor grads, Tensor gt, float eta, float eps, float gradScale, float clipNorm, Tensor normSq = nullptr);

// ---- beam search: n best continuations per sentence (kernels/nth_element.cu) ----------------------------------
// Replaces NthElement::getNBestList (src/translator/nth_element.cu:270-402; gMaxElement + gMaxElementUpdate, one
// iteration per returned element).  Both calls block until the result is on the host.
//
// NthElementRanges: scores is any tensor; range i covers the flat elements [rangeFirst[i], rangeFirst[i+1]) and
// returns cumN[i+1]-cumN[i] (value, flat index) pairs, best first; ties go to the lower index.
void NthElementRanges(Tensor scores, const std::vector<int>& rangeFirst, const std::vector<int>& cumN, std::vector<float>& outCosts,
                      std::vector<unsigned>& outKeys);
// NthElementLogSoftmax: the same result as
//   totalCosts = transpose(prevCosts + logsoftmax(logits), {2, 1, 0, 3});  getNBestList(n per sentence, totalCosts)
// read straight from the decoder's raw logits [beam, 1, batch, V]: row statistics and the row-local candidates come from
// ONE pass over the logits (the n best of a row do not depend on the row's normaliser), the merge per sentence adds
// prevCosts[row] + ((x - max) - log(sum)) in the LogSoftmax operator's order of operations.  Keys are those of the
// transposed tensor: (sentence * beam + hypothesis) * V + word.  first: only hypothesis 0 of every sentence competes.
// suppressWord >= 0: that word never wins (suppressUnk of src/translator/helpers.cu).
void NthElementLogSoftmax(Tensor logits, const std::vector<float>& prevCosts, int dimBatch, int beam, int n, bool first, int suppressWord,
                          std::vector<float>& outCosts, std::vector<unsigned>& outKeys);

}  // namespace marian

#include "kernels/element.h"
