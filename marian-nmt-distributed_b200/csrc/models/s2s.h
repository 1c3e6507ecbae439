// RNN encoder-decoder with attention ("s2s"; deep variants = Edinburgh/Nematus
// style stacks).  Layer composition, parameter prefixes and creation order
// follow the reference's src/models/s2s.h:7-371.
#pragma once

#include "models/encdec.h"
#include "rnn/rnn.h"

namespace marian {

class EncoderS2S : public EncoderBase {
public:
  EncoderS2S(Ptr<Options> options) : EncoderBase(options) {}

  // reference: s2s.h:9-99
  Expr applyEncoderRNN(Ptr<ExpressionGraph> graph, Expr embeddings, Expr mask, std::string type) {
    using namespace keywords;
    int first, second;
    if(type == "bidirectional" || type == "alternating") {
      // build two separate stacks, concatenate top outputs
      first = opt<int>("enc-depth");
      second = 0;
    } else {
      // one bidirectional layer, then unidirectional layers on top
      first = 1;
      second = opt<int>("enc-depth") - first;
    }

    auto forward = type == "alternating" ? rnn::dir::alternating_forward : rnn::dir::forward;
    auto backward = type == "alternating" ? rnn::dir::alternating_backward : rnn::dir::backward;

    float dropoutRnn = inference_ ? 0 : opt<float>("dropout-rnn");

    auto makeStack = [&](rnn::dir direction, const std::string& base) {
      auto r = rnn::rnn(graph)("type", opt<std::string>("enc-cell"))("direction", direction)(
          "dimInput", embeddings->shape()[-1])("dimState", opt<int>("dim-rnn"))("dropout", dropoutRnn)(
          "layer-normalization", opt<bool>("layer-normalization"))("skip", opt<bool>("skip"));
      for(int i = 1; i <= first; ++i) {
        auto stacked = rnn::stacked_cell(graph);
        for(int j = 1; j <= opt<int>("enc-cell-depth"); ++j) {
          std::string paramPrefix = base;
          if(i > 1)
            paramPrefix += "_l" + std::to_string(i);
          if(i > 1 || j > 1)
            paramPrefix += "_cell" + std::to_string(j);
          bool transition = (j > 1);
          stacked.push_back(rnn::cell(graph)("prefix", paramPrefix)("transition", transition));
        }
        r.push_back(stacked);
      }
      return r;
    };

    auto rnnFw = makeStack(forward, prefix_ + "_bi");
    auto rnnBw = makeStack(backward, prefix_ + "_bi_r");

    // NB: C++ leaves the evaluation order of the two transduce() calls inside
    // the reference's braced list well-defined (left to right): forward first.
    auto fw = rnnFw->transduce(embeddings, mask);
    auto bw = rnnBw->transduce(embeddings, mask);
    auto context = concatenate({fw, bw}, axis = -1);

    if(second > 0) {
      auto rnnUni = rnn::rnn(graph)("type", opt<std::string>("enc-cell"))("dimInput", 2 * opt<int>("dim-rnn"))(
          "dimState", opt<int>("dim-rnn"))("dropout", dropoutRnn)("layer-normalization",
                                                                    opt<bool>("layer-normalization"))("skip", opt<bool>("skip"));
      for(int i = first + 1; i <= second + first; ++i) {
        auto stacked = rnn::stacked_cell(graph);
        for(int j = 1; j <= opt<int>("enc-cell-depth"); ++j) {
          std::string paramPrefix = prefix_ + "_l" + std::to_string(i) + "_cell" + std::to_string(j);
          stacked.push_back(rnn::cell(graph)("prefix", paramPrefix));
        }
        rnnUni.push_back(stacked);
      }
      context = rnnUni->transduce(context);
    }
    return context;
  }

  Expr buildSourceEmbeddings(Ptr<ExpressionGraph> graph) {
    int dimVoc = opt<std::vector<int>>("dim-vocabs")[batchIndex_];
    int dimEmb = opt<int>("dim-emb");
    auto embFactory = embedding(graph)("dimVocab", dimVoc)("dimEmb", dimEmb);
    if(opt<bool>("tied-embeddings-src") || opt<bool>("tied-embeddings-all"))
      embFactory("prefix", "Wemb");
    else
      embFactory("prefix", prefix_ + "_Wemb");
    if(options_->has("embedding-fix-src"))
      embFactory("fixed", opt<bool>("embedding-fix-src"));
    return embFactory.construct();
  }

  virtual Ptr<EncoderState> build(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch) {
    using namespace keywords;
    auto embeddings = buildSourceEmbeddings(graph);

    Expr batchEmbeddings, batchMask;
    std::tie(batchEmbeddings, batchMask) = EncoderBase::lookup(embeddings, batch);

    float dropProb = inference_ ? 0 : opt<float>("dropout-src");
    if(dropProb) {
      int srcWords = batchEmbeddings->shape()[-3];
      auto dropMask = graph->dropout(dropProb, {srcWords, 1, 1});
      batchEmbeddings = dropout(batchEmbeddings, mask = dropMask);
    }

    Expr context = applyEncoderRNN(graph, batchEmbeddings, batchMask, opt<std::string>("enc-type"));
    return New<EncoderState>(context, batchMask, batch);
  }

  void clear() {}
};

class DecoderS2S : public DecoderBase {
private:
  Ptr<rnn::RNN> rnn_;

  // reference: s2s.h:178-234
  Ptr<rnn::RNN> constructDecoderRNN(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    float dropoutRnn = inference_ ? 0 : opt<float>("dropout-rnn");
    auto rnn = rnn::rnn(graph)("type", opt<std::string>("dec-cell"))("dimInput", opt<int>("dim-emb"))(
        "dimState", opt<int>("dim-rnn"))("dropout", dropoutRnn)("layer-normalization", opt<bool>("layer-normalization"))(
        "skip", opt<bool>("skip"));

    size_t decoderLayers = opt<size_t>("dec-depth");
    size_t decoderBaseDepth = opt<size_t>("dec-cell-base-depth");
    size_t decoderHighDepth = opt<size_t>("dec-cell-high-depth");

    // conditional GRU: cell1 -> attention -> cell2 [-> transition cells]
    auto baseCell = rnn::stacked_cell(graph);
    for(size_t i = 1; i <= decoderBaseDepth; ++i) {
      bool transition = (i > 2);
      auto paramPrefix = prefix_ + "_cell" + std::to_string(i);
      baseCell.push_back(rnn::cell(graph)("prefix", paramPrefix)("final", i > 1)("transition", transition));
      if(i == 1) {
        for(size_t k = 0; k < state->getEncoderStates().size(); ++k) {
          auto attPrefix = prefix_;
          if(state->getEncoderStates().size() > 1)
            attPrefix += "_att" + std::to_string(k + 1);
          auto encState = state->getEncoderStates()[k];
          baseCell.push_back(rnn::attention(graph)("prefix", attPrefix).set_state(encState));
        }
      }
    }
    rnn.push_back(baseCell);

    for(size_t i = 2; i <= decoderLayers; ++i) {
      auto highCell = rnn::stacked_cell(graph);
      for(size_t j = 1; j <= decoderHighDepth; j++) {
        auto paramPrefix = prefix_ + "_l" + std::to_string(i) + "_cell" + std::to_string(j);
        highCell.push_back(rnn::cell(graph)("prefix", paramPrefix));
      }
      rnn.push_back(highCell);
    }
    return rnn.construct();
  }

public:
  DecoderS2S(Ptr<Options> options) : DecoderBase(options) {}

  // reference: s2s.h:239-277
  virtual Ptr<DecoderState> startState(Ptr<ExpressionGraph> graph,
                                       Ptr<data::CorpusBatch> batch,
                                       std::vector<Ptr<EncoderState>>& encStates) {
    using namespace keywords;

    std::vector<Expr> meanContexts;
    for(auto& encState : encStates)
      meanContexts.push_back(weighted_average(encState->getContext(), encState->getMask(), axis = -3));

    Expr start;
    if(!meanContexts.empty()) {
      auto mlp = mlp::mlp(graph).push_back(mlp::dense(graph)("prefix", prefix_ + "_ff_state")("dim", opt<int>("dim-rnn"))(
          "activation", mlp::act::tanh)("layer-normalization", opt<bool>("layer-normalization")));
      auto built = mlp.construct();
      start = meanContexts.size() == 1 ? built->apply(meanContexts[0]) : Expr();
      ABORT_IF(!start, "multi-encoder start state is not supported");
    } else {
      int dimBatch = (int)batch->size();
      int dimRnn = opt<int>("dim-rnn");
      start = graph->constant({dimBatch, dimRnn}, init = inits::zeros);
    }

    rnn::States startStates(opt<size_t>("dec-depth"), {start, start});
    return New<DecoderState>(startStates, nullptr, encStates);
  }

  // reference: s2s.h:279-361
  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    using namespace keywords;

    auto embeddings = state->getTargetEmbeddings();

    float dropoutTrg = inference_ ? 0 : opt<float>("dropout-trg");
    if(dropoutTrg) {
      int trgWords = embeddings->shape()[-3];
      auto trgWordDrop = graph->dropout(dropoutTrg, {trgWords, 1, 1});
      embeddings = dropout(embeddings, mask = trgWordDrop);
    }

    if(!rnn_)
      rnn_ = constructDecoderRNN(graph, state);

    auto decoderContext = rnn_->transduce(embeddings, state->getStates());
    rnn::States decoderStates = rnn_->lastCellStates();

    std::vector<Expr> alignedContexts;
    for(size_t k = 0; k < state->getEncoderStates().size(); ++k) {
      auto att = rnn_->at(0)->as<rnn::StackedCell>()->at((int)k + 1)->as<rnn::Attention>();
      alignedContexts.push_back(att->getContext());
    }

    Expr alignedContext;
    if(alignedContexts.size() > 1)
      alignedContext = concatenate(alignedContexts, axis = -1);
    else if(alignedContexts.size() == 1)
      alignedContext = alignedContexts[0];

    auto layer1 = mlp::dense(graph)("prefix", prefix_ + "_ff_logit_l1")("dim", opt<int>("dim-emb"))(
        "activation", mlp::act::tanh)("layer-normalization", opt<bool>("layer-normalization"));

    int dimTrgVoc = opt<std::vector<int>>("dim-vocabs")[batchIndex_];
    auto layer2 = mlp::dense(graph)("prefix", prefix_ + "_ff_logit_l2")("dim", dimTrgVoc);
    if(opt<bool>("tied-embeddings") || opt<bool>("tied-embeddings-all")) {
      std::string tiedPrefix = prefix_ + "_Wemb";
      if(opt<bool>("tied-embeddings-all") || opt<bool>("tied-embeddings-src"))
        tiedPrefix = "Wemb";
      layer2.tie_transposed("W", tiedPrefix);
    }

    auto output = mlp::mlp(graph).push_back(layer1).push_back(layer2);

    Expr logits;
    if(alignedContext)
      logits = output->apply(embeddings, decoderContext, alignedContext);
    else
      logits = output->apply(embeddings, decoderContext);

    return New<DecoderState>(decoderStates, logits, state->getEncoderStates());
  }

  virtual const std::vector<Expr> getAlignments(int i = 0) {
    auto att = rnn_->at(0)->as<rnn::StackedCell>()->at(i + 1)->as<rnn::Attention>();
    return att->getAlignments();
  }

  void clear() { rnn_ = nullptr; }
};

}  // namespace marian
