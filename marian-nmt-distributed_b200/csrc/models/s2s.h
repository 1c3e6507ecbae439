// RNN encoder-decoder with attention ("s2s"; the deep variants are the Edinburgh / Nematus style
// stacks).  Behavioural contract = the reference's src/models/s2s.h:7-371: the same layer
// composition, parameter prefixes and parameter creation order (the order fixes the random
// initialisation stream).  Organisation is this repo's own: stack recipes are written as small
// tables (S2SRecipes) that one routine turns into rnn:: factories.
#pragma once

#include "models/encdec.h"
#include "rnn/rnn.h"

namespace marian {

// Shared helpers of encoder and decoder.
struct S2SRecipes {
  // factory of one RNN with the model-wide cell options
  static rnn::rnn stackFactory(Ptr<ExpressionGraph> graph, Ptr<Options> o, const std::string& cellKey, int dimInput, float dropProb) {
    auto f = rnn::rnn(graph);
    f("type", o->get<std::string>(cellKey));
    f("dimInput", dimInput);
    f("dimState", o->get<int>("dim-rnn"));
    f("dropout", dropProb);
    f("layer-normalization", o->get<bool>("layer-normalization"));
    f("skip", o->get<bool>("skip"));
    return f;
  }

  // word-level dropout of an embedding sequence [T, B, d] (one Bernoulli per word)
  static Expr dropWords(Ptr<ExpressionGraph> graph, Expr sequence, float prob) {
    if(!prob)
      return sequence;
    int words = sequence->shape()[-3];
    return dropout(sequence, keywords::mask = graph->dropout(prob, {words, 1, 1}));
  }

  static std::string embeddingName(Ptr<Options> o, const std::string& own, bool sourceSide) {
    bool all = o->get<bool>("tied-embeddings-all");
    bool shared = sourceSide ? (o->get<bool>("tied-embeddings-src") || all) : (all || o->get<bool>("tied-embeddings-src"));
    return shared ? std::string("Wemb") : own + "_Wemb";
  }
};

class EncoderS2S : public EncoderBase {
public:
  EncoderS2S(Ptr<Options> options) : EncoderBase(options) {}

  // enc-type "bidirectional" / "alternating": two full-depth stacks run in opposite directions,
  // top outputs concatenated.  Otherwise ("bi-unidirectional"): ONE bidirectional layer, then
  // enc-depth - 1 unidirectional layers over the concatenation.        (reference :9-99)
  Expr encode(Ptr<ExpressionGraph> graph, Expr embeddings, Expr mask) {
    using namespace keywords;
    const std::string kind = opt<std::string>("enc-type");
    const bool twoStacks = kind == "bidirectional" || kind == "alternating";
    const int depth = opt<int>("enc-depth");
    const int biLayers = twoStacks ? depth : 1;
    const int cellDepth = opt<int>("enc-cell-depth");
    const float dropProb = inference_ ? 0 : opt<float>("dropout-rnn");

    // cell names of layer i, position j of a directional stack rooted at `base`
    auto cellName = [](const std::string& base, int layer, int pos) {
      std::string n = base;
      if(layer > 1)
        n += "_l" + std::to_string(layer);
      if(layer > 1 || pos > 1)
        n += "_cell" + std::to_string(pos);
      return n;
    };
    auto directional = [&](rnn::dir direction, const std::string& base) {
      auto f = S2SRecipes::stackFactory(graph, options_, "enc-cell", embeddings->shape()[-1], dropProb);
      f("direction", direction);
      for(int layer = 1; layer <= biLayers; ++layer) {
        auto cells = rnn::stacked_cell(graph);
        for(int pos = 1; pos <= cellDepth; ++pos)
          cells.push_back(rnn::cell(graph)("prefix", cellName(base, layer, pos))("transition", pos > 1));
        f.push_back(cells);
      }
      return f;
    };

    const bool alternate = kind == "alternating";
    auto left = directional(alternate ? rnn::dir::alternating_forward : rnn::dir::forward, prefix_ + "_bi");
    auto right = directional(alternate ? rnn::dir::alternating_backward : rnn::dir::backward, prefix_ + "_bi_r");
    // forward stack first: its parameters are created (and seeded) before the backward stack's
    Expr fw = left->transduce(embeddings, mask);
    // the two directions share nothing but their input: the backward-direction stack is a chain (lane) of its own,
    // forward and backward passes run both stacks side by side (ExpressionGraph::setLane, tensors/device.h)
    graph->setLane(options_->get<bool>("rnn-lanes", true) ? 1 : 0);
    Expr bw = right->transduce(embeddings, mask);
    graph->setLane(0);
    Expr context = concatenate({fw, bw}, axis = -1);

    if(!twoStacks && depth > 1) {
      auto upper = S2SRecipes::stackFactory(graph, options_, "enc-cell", 2 * opt<int>("dim-rnn"), dropProb);
      for(int layer = 2; layer <= depth; ++layer) {
        auto cells = rnn::stacked_cell(graph);
        for(int pos = 1; pos <= cellDepth; ++pos)
          cells.push_back(rnn::cell(graph)("prefix", prefix_ + "_l" + std::to_string(layer) + "_cell" + std::to_string(pos)));
        upper.push_back(cells);
      }
      context = upper->transduce(context);
    }
    return context;
  }

  virtual Ptr<EncoderState> build(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch) {
    auto table = embedding(graph)("dimVocab", opt<std::vector<int>>("dim-vocabs")[batchIndex_])("dimEmb", opt<int>("dim-emb"));
    table("prefix", S2SRecipes::embeddingName(options_, prefix_, true));
    if(options_->has("embedding-fix-src"))
      table("fixed", opt<bool>("embedding-fix-src"));

    Expr words, padding;
    std::tie(words, padding) = EncoderBase::lookup(table.construct(), batch);
    words = S2SRecipes::dropWords(graph, words, inference_ ? 0 : opt<float>("dropout-src"));
    return New<EncoderState>(encode(graph, words, padding), padding, batch);
  }

  void clear() {}
};

class DecoderS2S : public DecoderBase {
private:
  Ptr<rnn::RNN> rnn_;

  // Layer 1 is the conditional cell: cell1 -> one attention per encoder -> cell2 [-> transition
  // cells]; layers 2.. are plain stacks of dec-cell-high-depth cells.     (reference :178-234)
  Ptr<rnn::RNN> makeRnn(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    auto f = S2SRecipes::stackFactory(graph, options_, "dec-cell", opt<int>("dim-emb"), inference_ ? 0 : opt<float>("dropout-rnn"));
    const auto& encoders = state->getEncoderStates();

    auto conditional = rnn::stacked_cell(graph);
    const size_t baseDepth = opt<size_t>("dec-cell-base-depth");
    for(size_t pos = 1; pos <= baseDepth; ++pos) {
      conditional.push_back(rnn::cell(graph)("prefix", prefix_ + "_cell" + std::to_string(pos))("final", pos > 1)("transition", pos > 2));
      if(pos != 1)
        continue;
      for(size_t e = 0; e < encoders.size(); ++e) {
        std::string name = encoders.size() > 1 ? prefix_ + "_att" + std::to_string(e + 1) : prefix_;
        conditional.push_back(rnn::attention(graph)("prefix", name).set_state(encoders[e]));
      }
    }
    f.push_back(conditional);

    const size_t layers = opt<size_t>("dec-depth"), highDepth = opt<size_t>("dec-cell-high-depth");
    for(size_t layer = 2; layer <= layers; ++layer) {
      auto cells = rnn::stacked_cell(graph);
      for(size_t pos = 1; pos <= highDepth; ++pos)
        cells.push_back(rnn::cell(graph)("prefix", prefix_ + "_l" + std::to_string(layer) + "_cell" + std::to_string(pos)));
      f.push_back(cells);
    }
    return f.construct();
  }

  Ptr<rnn::Attention> attentionOf(size_t encoder) { return rnn_->at(0)->as<rnn::StackedCell>()->at((int)encoder + 1)->as<rnn::Attention>(); }

public:
  DecoderS2S(Ptr<Options> options) : DecoderBase(options) {}

  // start state = tanh(W mean_t(context) + b) for every layer (zeros without an encoder)   (:239-277)
  virtual Ptr<DecoderState> startState(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch, std::vector<Ptr<EncoderState>>& encStates) {
    using namespace keywords;
    Expr start;
    if(encStates.empty()) {
      start = graph->constant({(int)batch->size(), opt<int>("dim-rnn")}, init = inits::zeros);
    } else {
      ABORT_IF(encStates.size() != 1, "multi-encoder start state is not supported");
      Expr meanContext = weighted_average(encStates[0]->getContext(), encStates[0]->getMask(), axis = -3);
      auto bridge = mlp::dense(graph)("prefix", prefix_ + "_ff_state")("dim", opt<int>("dim-rnn"))("activation", mlp::act::tanh)(
          "layer-normalization", opt<bool>("layer-normalization"));
      start = mlp::mlp(graph).push_back(bridge).construct()->apply(meanContext);
    }
    return New<DecoderState>(rnn::States(opt<size_t>("dec-depth"), {start, start}), nullptr, encStates);
  }

  // all target positions at once in training; logits = W2 tanh(W1 [embedding; state; context])   (:279-361)
  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    using namespace keywords;
    Expr words = S2SRecipes::dropWords(graph, state->getTargetEmbeddings(), inference_ ? 0 : opt<float>("dropout-trg"));
    if(!rnn_)
      rnn_ = makeRnn(graph, state);

    Expr states = rnn_->transduce(words, state->getStates());
    rnn::States carried = rnn_->lastCellStates();

    std::vector<Expr> contexts;
    for(size_t e = 0; e < state->getEncoderStates().size(); ++e)
      contexts.push_back(attentionOf(e)->getContext());

    auto hidden = mlp::dense(graph)("prefix", prefix_ + "_ff_logit_l1")("dim", opt<int>("dim-emb"))("activation", mlp::act::tanh)(
        "layer-normalization", opt<bool>("layer-normalization"));
    auto vocabulary = mlp::dense(graph)("prefix", prefix_ + "_ff_logit_l2")("dim", opt<std::vector<int>>("dim-vocabs")[batchIndex_]);
    if(opt<bool>("tied-embeddings") || opt<bool>("tied-embeddings-all"))
      vocabulary.tie_transposed("W", S2SRecipes::embeddingName(options_, prefix_, false));
    auto readout = mlp::mlp(graph).push_back(hidden).push_back(vocabulary);

    Expr logits;
    if(contexts.empty())
      logits = readout->apply(words, states);
    else
      logits = readout->apply(words, states, contexts.size() == 1 ? contexts[0] : concatenate(contexts, axis = -1));
    return New<DecoderState>(carried, logits, state->getEncoderStates());
  }

  virtual const std::vector<Expr> getAlignments(int i = 0) { return attentionOf((size_t)i)->getAlignments(); }

  void clear() { rnn_ = nullptr; }
};

}  // namespace marian
