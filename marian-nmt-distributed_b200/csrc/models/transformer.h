// Transformer encoder / decoder (Vaswani et al. 2017) on the operator API.
//
// Behavioural contract = the reference's src/models/transformer.h:7-662: same parameter names
// and the same parameter CREATION ORDER (it fixes the random-initialisation stream, one seed
// increment per tensor), embeddings * sqrt(d) + sinusoid table, batch-major layers, additive
// -99999999 masks, sub-layer recipe strings ("dan": dropout, add residual, layer-norm(1e-6)),
// swish feed-forward, vocabulary projection through mlp::dense.
//
// Organisation is this repo's own: a stateless TransformerSublayers toolbox parameterised by a
// small Env (graph, options, inference flag), one routine for pre- and post-processing recipes,
// and an attention core that is ONE fused operator (multi_head_attention, kernels/attention.cu)
// whenever the shape allows it; the unfused node sequence of the reference (split heads, bdot,
// mask add, softmax, bdot, join heads) stays available as "transformer-fused-attention=false"
// and is what the CPU oracle builds.
#pragma once

#include <cmath>

#include "models/encdec.h"

namespace marian {

struct TransformerSublayers {
  struct Env {
    Ptr<ExpressionGraph> graph;
    Ptr<Options> options;
    bool inference;

    float dropout(const char* key) const { return inference ? 0.f : options->get<float>(key); }
    std::string recipe(const char* key) const { return options->get<std::string>(key); }
    bool fuseAttention() const {
      return options->get<bool>("transformer-fused-attention", std::string(device::backendName()) == "cuda");
    }
    // key/value projections on the side stream while the query projection runs (tf32 / fp32
    // GEMM modes only: the packed-operand modes share one scratch cache between streams)
    bool concurrentProjections() const {
      auto mode = getGemmMode(graph->getBackend()->getGemmHandle());
      return std::string(device::backendName()) == "cuda" && (mode == GemmMode::TF32 || mode == GemmMode::FP32)
             && options->get<bool>("transformer-concurrent-projections", true);
    }
    bool fuseResidualNorm() const {
      return options->get<bool>("transformer-fused-residual-norm", std::string(device::backendName()) == "cuda");
    }
  };

  // [.., T, B, d] <-> [.., B, T, d]
  static Expr swapTimeBatch(Expr x) { return transpose(x, {0, 2, 1, 3}); }

  // x + sinusoid(position): first half of the channels sin, second half cos, geometric
  // wavelengths 1 .. 10000 (float arithmetic as in the reference, transformer.h:11-35)
  static Expr addPositions(const Env& env, Expr x, int firstPosition = 0) {
    const int channels = x->shape()[-1];
    const int steps = x->shape()[-3];
    const int half = channels / 2;
    const float logStep = std::log(10000.f) / ((float)half - 1.f);

    std::vector<float> table((size_t)steps * channels, 0.f);
    for(int t = 0; t < steps; ++t) {
      float* row = table.data() + (size_t)t * channels;
      const int position = firstPosition + t;
      for(int c = 0; c < half; ++c) {
        float angle = position * std::exp(c * -logStep);
        row[c] = std::sin(angle);
        row[half + c] = std::cos(angle);
      }
    }
    return x + env.graph->constant({steps, 1, channels}, keywords::init = inits::from_vector(table));
  }

  // lower-triangular 0/1 matrix [1, n, n]: position i may look at j <= i
  static Expr causalMask(const Env& env, int n) {
    std::vector<float> tri((size_t)n * n, 0.f);
    for(int i = 0; i < n; ++i)
      std::fill(tri.begin() + (size_t)i * n, tri.begin() + (size_t)i * n + i + 1, 1.f);
    return env.graph->constant({1, n, n}, keywords::init = inits::from_vector(tri));
  }

  // 0/1 mask [.., B, rows, T] -> additive {0, -99999999} mask [B, 1, rows, T] (broadcast over heads)
  static Expr additiveMask(Expr keep) {
    auto s = keep->shape();
    return reshape((1 - keep) * -99999999.f, {s[-3], 1, s[-2], s[-1]});
  }

  // [1, B, T, 1] time-major padding mask -> [1, B, 1, T]
  static Expr keyMask(Expr paddingMask, int batch, int steps) {
    return reshape(swapTimeBatch(atleast_nd(paddingMask, 4)), {1, batch, 1, steps});
  }

  // One recipe interpreter for both ends of a sub-layer.  Letters: d dropout, a add the residual,
  // h highway gate with the residual, n layer-norm.  `residual` == nullptr is the pre-processing
  // flavour (no a/h; layer-norm parameters carry the "_pre" suffix).
  static Expr runRecipe(const Env& env, const std::string& prefix, const std::string& recipe, Expr x, Expr residual, float dropProb) {
    using namespace keywords;
    const int width = x->shape()[-1];
    const std::string suffix = residual ? "" : "_pre";
    for(size_t pos = 0; pos < recipe.size(); ++pos) {
      const char step = recipe[pos];
      // "an": add the residual and normalise in ONE operator (no parameters are created by "a",
      // so the creation order of the layer-norm parameters is unchanged)
      if(step == 'a' && residual && pos + 1 < recipe.size() && recipe[pos + 1] == 'n' && env.fuseResidualNorm() && LayerNormResidualFusable(width)
         && x->shape() == residual->shape()) {
        auto gain = env.graph->param(prefix + "_ln_scale" + suffix, {1, width}, init = inits::ones);
        auto shift = env.graph->param(prefix + "_ln_bias" + suffix, {1, width}, init = inits::zeros);
        x = residual_layer_norm(x, residual, gain, shift, 1e-6);
        ++pos;
        continue;
      }
      switch(step) {
        case 'd':
          if(dropProb > 0.f)
            x = dropout(x, mask = env.graph->dropout(dropProb, x->shape()));
          break;
        case 'a':
          if(residual)
            x = x + residual;
          break;
        case 'h':
          if(residual) {
            auto Wh = env.graph->param(prefix + "_Wh", {width, width}, init = inits::glorot_uniform);
            auto bh = env.graph->param(prefix + "_bh", {1, width}, init = inits::zeros);
            x = highway(x, residual, affine(residual, Wh, bh));
          }
          break;
        case 'n': {
          auto gain = env.graph->param(prefix + "_ln_scale" + suffix, {1, width}, init = inits::ones);
          auto shift = env.graph->param(prefix + "_ln_bias" + suffix, {1, width}, init = inits::zeros);
          x = layer_norm(x, gain, shift, 1e-6);
          break;
        }
        default: break;
      }
    }
    return x;
  }

  static Expr linear(const Env& env, const std::string& weight, const std::string& bias, Expr x, int in, int out) {
    using namespace keywords;
    auto W = env.graph->param(weight, {in, out}, init = inits::glorot_uniform);
    auto b = env.graph->param(bias, {1, out}, init = inits::zeros);
    return affine(x, W, b);
  }

  // ---- unfused attention core (the reference's node sequence, transformer.h:58-77,153-192) ----
  static Expr splitHeads(Expr x, int heads) {
    auto s = x->shape();
    return transpose(reshape(x, {s[-3] * s[-4], s[-2], heads, s[-1] / heads}), {0, 2, 1, 3});
  }
  static Expr joinHeads(Expr x, int beam) {
    auto s = x->shape();  // [B*beam, heads, T, dk]
    return reshape(transpose(x, {0, 2, 1, 3}), {beam, s[-4] / beam, s[-2], s[-3] * s[-1]});
  }
  static Expr attendUnfused(const Env& env, Expr q, Expr k, Expr v, Expr mask, int heads, int beam) {
    using namespace keywords;
    Expr qh = splitHeads(q, heads), kh = splitHeads(k, heads), vh = splitHeads(v, heads);
    const float scale = 1.0f / std::sqrt((float)kh->shape()[-1]);
    const int beamRatio = qh->shape()[-4] / kh->shape()[-4];
    if(beamRatio > 1) {  // beam search: keys/values are shared by the hypotheses of a sentence
      kh = repeat(kh, beamRatio, axis = -4);
      vh = repeat(vh, beamRatio, axis = -4);
    }
    Expr weights = softmax(bdot(qh, kh, false, true, scale) + mask);
    float dropProb = env.dropout("transformer-dropout-attention");
    if(dropProb)
      weights = dropout(weights, keywords::mask = env.graph->dropout(dropProb, weights->shape()));
    return joinHeads(bdot(weights, vh), beam);
  }

  // softmax(q k^T / sqrt(dk) + mask) v over `heads` heads; q, k, v are [beam, B, T, d] projections
  static Expr attend(const Env& env, Expr q, Expr k, Expr v, Expr mask, int heads) {
    const int beam = q->shape()[-4];
    const int width = q->shape()[-1];
    bool fusable = env.fuseAttention() && beam == 1 && k->shape()[-4] == 1 && mask && env.dropout("transformer-dropout-attention") == 0.f
                   && AttentionFusable(q->shape()[-2], k->shape()[-2], width, heads);
    if(fusable) {  // the fused operator wants one mask row per key (B*Tk) or per query and key (B*Tq*Tk); decoding steps
                   // carry a broadcast [1,1,1,1] causal mask and take the node sequence
      const size_t sentences = q->shape().elements() / ((size_t)q->shape()[-2] * width);
      const size_t keys = k->shape()[-2], queries = q->shape()[-2], have = mask->shape().elements();
      fusable = have == sentences * keys || have == sentences * queries * keys;
    }
    if(fusable)
      return multi_head_attention(q, k, v, mask, heads, 1.0f / std::sqrt((float)(width / heads)));
    return attendUnfused(env, q, k, v, mask, heads, beam);
  }

  // projections + attention core + output projection.  Parameter order: Wq bq, then per memory
  // Wk bk Wv bv, then Wo bo (reference :194-261).  Several memories (multi-source) are attended
  // separately and concatenated before Wo.
  static Expr multiHead(const Env& env, const std::string& prefix, int heads, Expr query, const std::vector<Expr>& memories, const std::vector<Expr>& masks) {
    using namespace keywords;
    const int width = query->shape()[-1];
    // parameters first, in the reference's creation order (it fixes the initialisation stream) ...
    auto Wq = env.graph->param(prefix + "_Wq", {width, width}, init = inits::glorot_uniform);
    auto bq = env.graph->param(prefix + "_bq", {1, width}, init = inits::zeros);
    struct KV {
      Expr Wk, bk, Wv, bv;
    };
    std::vector<KV> kv;
    for(size_t m = 0; m < memories.size(); ++m) {
      std::string p = m == 0 ? prefix : prefix + "_enc" + std::to_string(m + 1);
      KV w;
      w.Wk = env.graph->param(p + "_Wk", {width, width}, init = inits::glorot_uniform);
      w.bk = env.graph->param(p + "_bk", {1, width}, init = inits::zeros);
      w.Wv = env.graph->param(p + "_Wv", {width, width}, init = inits::glorot_uniform);
      w.bv = env.graph->param(p + "_bv", {1, width}, init = inits::zeros);
      kv.push_back(w);
    }
    // ... then the products: keys and values go to the side stream, the query projection runs
    // concurrently on the main stream, the attention core joins them
    const bool concurrent = env.concurrentProjections();
    std::vector<Expr> keys, values;
    for(size_t m = 0; m < memories.size(); ++m) {
      Expr k = affine(memories[m], kv[m].Wk, kv[m].bk);
      Expr v = affine(memories[m], kv[m].Wv, kv[m].bv);
      k->setConcurrent(concurrent);
      v->setConcurrent(concurrent);
      keys.push_back(k);
      values.push_back(v);
    }
    Expr q = affine(query, Wq, bq);

    std::vector<Expr> contexts;
    for(size_t m = 0; m < memories.size(); ++m)
      contexts.push_back(attend(env, q, keys[m], values[m], masks[m], heads));
    Expr joined = contexts.size() == 1 ? contexts.front() : concatenate(contexts, keywords::axis = -1);
    return linear(env, prefix + "_Wo", prefix + "_bo", joined, joined->shape()[-1], width);
  }

  // pre-recipe -> multi-head attention over `memories` -> post-recipe with the residual
  static Expr attentionSublayer(const Env& env, const std::string& prefix, Expr x, const std::vector<Expr>& memories, const std::vector<Expr>& masks) {
    const float dropProb = env.dropout("transformer-dropout");
    Expr h = runRecipe(env, prefix + "_Wo", env.recipe("transformer-preprocess"), x, nullptr, dropProb);
    h = multiHead(env, prefix, (int)env.options->get<float>("transformer-heads"), h, memories, masks);
    return runRecipe(env, prefix + "_Wo", env.recipe("transformer-postprocess"), h, x, dropProb);
  }

  // pre-recipe -> W2 swish(W1 x + b1) + b2 -> post-recipe with the residual (reference :318-350)
  static Expr feedForwardSublayer(const Env& env, const std::string& prefix, Expr x) {
    const int width = x->shape()[-1];
    const int inner = env.options->get<int>("transformer-dim-ffn");
    const float dropProb = env.dropout("transformer-dropout");
    Expr h = runRecipe(env, prefix + "_ffn", env.recipe("transformer-preprocess"), x, nullptr, dropProb);
    // all four parameters are created before the first product (initialisation stream order)
    using namespace keywords;
    auto W1 = env.graph->param(prefix + "_W1", {width, inner}, init = inits::glorot_uniform);
    auto b1 = env.graph->param(prefix + "_b1", {1, inner}, init = inits::zeros);
    auto W2 = env.graph->param(prefix + "_W2", {inner, width}, init = inits::glorot_uniform);
    auto b2 = env.graph->param(prefix + "_b2", {1, width}, init = inits::zeros);
    h = affine(swish(affine(h, W1, b1)), W2, b2);
    return runRecipe(env, prefix + "_ffn", env.recipe("transformer-postprocess"), h, x, dropProb);
  }

  // sqrt(d) * embeddings + positions, batch-major, embedding recipe applied
  static Expr embedInput(const Env& env, const std::string& prefix, Expr embeddings, float wordDropout, int firstPosition) {
    using namespace keywords;
    if(wordDropout) {
      int words = embeddings->shape()[-3];
      embeddings = dropout(embeddings, mask = env.graph->dropout(wordDropout, {words, 1, 1}));
    }
    const int width = embeddings->shape()[-1];
    Expr x = addPositions(env, std::sqrt((float)width) * embeddings, firstPosition);
    x = swapTimeBatch(atleast_nd(x, 4));
    return runRecipe(env, prefix + "_emb", env.recipe("transformer-postprocess-emb"), x, nullptr, env.dropout("transformer-dropout"));
  }
};

class EncoderTransformer : public EncoderBase {
public:
  EncoderTransformer(Ptr<Options> options) : EncoderBase(options) {}

  Expr sourceEmbeddings(Ptr<ExpressionGraph> graph) {
    auto factory = embedding(graph)("dimVocab", opt<std::vector<int>>("dim-vocabs")[batchIndex_])("dimEmb", opt<int>("dim-emb"));
    bool shared = opt<bool>("tied-embeddings-src") || opt<bool>("tied-embeddings-all");
    factory("prefix", shared ? std::string("Wemb") : prefix_ + "_Wemb");
    if(options_->has("embedding-fix-src"))
      factory("fixed", opt<bool>("embedding-fix-src"));
    return factory.construct();
  }

  // reference: transformer.h:384-449
  Ptr<EncoderState> build(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch) {
    typedef TransformerSublayers T;
    T::Env env{graph, options_, inference_};
    const int sentences = (int)batch->size();
    const int steps = (int)(*batch)[batchIndex_]->batchWidth();

    Expr tokens, padding;
    std::tie(tokens, padding) = EncoderBase::lookup(sourceEmbeddings(graph), batch);

    Expr x = T::embedInput(env, prefix_, tokens, env.dropout("dropout-src"), 0);
    Expr padding4 = atleast_nd(padding, 4);
    Expr mask = T::additiveMask(T::keyMask(padding4, sentences, steps));

    const int depth = opt<int>("enc-depth");
    for(int l = 1; l <= depth; ++l) {
      std::string layer = prefix_ + "_l" + std::to_string(l);
      x = T::attentionSublayer(env, layer + "_self", x, {x}, {mask});
      x = T::feedForwardSublayer(env, layer + "_ffn", x);
    }
    return New<EncoderState>(T::swapTimeBatch(x), padding4, batch);
  }

  void clear() {}
};

class TransformerState : public DecoderState {
public:
  TransformerState(const rnn::States& states, Expr probs, std::vector<Ptr<EncoderState>>& encStates)
      : DecoderState(states, probs, encStates) {}

  // The cached layer inputs [beam, batch, time, d] of the chosen hypotheses: hypothesis h owns rows
  // h*time .. h*time+time-1 of the flattened cache.   reference: transformer.h:461-480
  virtual Ptr<DecoderState> select(const std::vector<size_t>& selIdx, int beamSize) {
    const int width = states_[0].output->shape()[-1];
    const int steps = states_[0].output->shape()[-2];
    const int sentences = (int)selIdx.size() / beamSize;
    std::vector<size_t> cacheRows;
    cacheRows.reserve(selIdx.size() * steps);
    for(size_t h : selIdx)
      for(int t = 0; t < steps; ++t)
        cacheRows.push_back(h * steps + t);
    rnn::States picked;
    for(auto& layer : states_)
      picked.push_back({reshape(rows(flatten_2d(layer.output), cacheRows), {beamSize, sentences, steps, width}), nullptr});
    return New<TransformerState>(picked, probs_, encStates_);
  }
};

class DecoderTransformer : public DecoderBase {
public:
  DecoderTransformer(Ptr<Options> options) : DecoderBase(options) {}

  virtual Ptr<DecoderState> startState(Ptr<ExpressionGraph>, Ptr<data::CorpusBatch>, std::vector<Ptr<EncoderState>>& encStates) {
    return New<TransformerState>(rnn::States(), nullptr, encStates);
  }

  // reference: transformer.h:495-662.  During training the whole target is one "step".
  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    typedef TransformerSublayers T;
    using namespace keywords;
    T::Env env{graph, options_, inference_};

    // layer inputs cached by previous decoding steps: new queries attend to old + new positions
    auto history = state->getStates();
    const int firstPosition = (history.size() == 0) ? 0 : history[0].output->shape()[-2];

    Expr x = T::embedInput(env, prefix_, state->getTargetEmbeddings(), env.dropout("dropout-trg"), firstPosition);
    const int beam = x->shape()[-4];
    const int sentences = x->shape()[-3];
    const int steps = x->shape()[-2];

    Expr selfKeep = T::causalMask(env, steps);
    if(Expr padding = state->getTargetMask())
      selfKeep = selfKeep * T::keyMask(padding, sentences, steps);
    Expr selfMask = T::additiveMask(selfKeep);

    std::vector<Expr> memories, memoryMasks;
    for(auto enc : state->getEncoderStates()) {
      Expr memory = T::swapTimeBatch(enc->getContext());
      Expr m = T::additiveMask(T::keyMask(enc->getMask(), sentences, memory->shape()[-2]));
      if(beam > 1)
        m = repeat(m, beam, axis = -4);
      memories.push_back(memory);
      memoryMasks.push_back(m);
    }

    rnn::States cache;
    const int depth = opt<int>("dec-depth");
    for(int l = 1; l <= depth; ++l) {
      std::string layer = prefix_ + "_l" + std::to_string(l);
      Expr selfMemory = (history.size() == 0) ? x : concatenate({history[l - 1].output, x}, axis = -2);
      cache.push_back({selfMemory, nullptr});

      x = T::attentionSublayer(env, layer + "_self", x, {selfMemory}, {selfMask});
      // one cross-attention sub-layer per encoder, stacked
      for(size_t e = 0; e < memories.size(); ++e) {
        std::string name = layer + "_context" + (e == 0 ? std::string() : "_enc" + std::to_string(e + 1));
        x = T::attentionSublayer(env, name, x, {memories[e]}, {memoryMasks[e]});
      }
      x = T::feedForwardSublayer(env, layer + "_ffn", x);
    }

    auto projection = mlp::dense(graph)("prefix", prefix_ + "_ff_logit_out")("dim", opt<std::vector<int>>("dim-vocabs").back());
    if(opt<bool>("tied-embeddings") || opt<bool>("tied-embeddings-all")) {
      bool shared = opt<bool>("tied-embeddings-all") || opt<bool>("tied-embeddings-src");
      projection.tie_transposed("W", shared ? std::string("Wemb") : prefix_ + "_Wemb");
    }
    Expr logits = mlp::mlp(graph).push_back(projection)->apply(T::swapTimeBatch(x));
    return New<TransformerState>(cache, logits, state->getEncoderStates());
  }

  void clear() {}
};

}  // namespace marian
