// Encoder-decoder skeleton: embedding lookup, ground-truth preparation, cost.
//
// Node construction order and formulas follow the reference
// (src/models/encdec.h:14-33 lookup, :74-123 groundTruth, :334-363 build).
// Batch-dependent constants (indices, masks, labels) are created through
// ExpressionGraph::batchConstant / rows(.., fill, batch) so a captured step can
// be replayed with the next batch (training/graph_replay.h).
#pragma once

#include "common/options.h"
#include "data/batch.h"
#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "layers/generic.h"
#include "models/states.h"

namespace marian {

namespace models {
class ModelBase {
public:
  virtual ~ModelBase() {}
  virtual Expr build(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch, bool clearGraph = true) = 0;
  virtual void clear(Ptr<ExpressionGraph> graph) = 0;
};
}  // namespace models

class EncoderBase {
protected:
  Ptr<Options> options_;
  std::string prefix_{"encoder"};
  bool inference_{false};
  size_t batchIndex_{0};

  virtual std::tuple<Expr, Expr> lookup(Expr srcEmbeddings, Ptr<data::CorpusBatch> batch) {
    using namespace keywords;
    size_t bi = batchIndex_;
    auto subBatch = (*batch)[bi];
    int dimBatch = (int)subBatch->batchSize();
    int dimEmb = srcEmbeddings->shape()[-1];
    int dimWords = (int)subBatch->batchWidth();

    auto graph = srcEmbeddings->graph();
    auto chosenEmbeddings = rows(
        srcEmbeddings,
        subBatch->indices(),
        [bi](const data::CorpusBatch& b, int* dst) {
          auto& idx = b[bi]->indices();
          for(size_t i = 0; i < idx.size(); ++i)
            dst[i] = (int)idx[i];
        },
        batch);
    auto batchEmbeddings = reshape(chosenEmbeddings, {dimWords, dimBatch, dimEmb});
    auto batchMask = graph->batchConstant(
        {dimWords, dimBatch, 1},
        [bi](const data::CorpusBatch& b, float* dst) {
          auto& m = b[bi]->mask();
          std::copy(m.begin(), m.end(), dst);
        },
        batch);
    return std::make_tuple(batchEmbeddings, batchMask);
  }

public:
  EncoderBase(Ptr<Options> options)
      : options_(options),
        prefix_(options->get<std::string>("prefix", "encoder")),
        inference_(options->get<bool>("inference", false)),
        batchIndex_(options->get<size_t>("index", 0)) {}
  virtual ~EncoderBase() {}

  virtual Ptr<EncoderState> build(Ptr<ExpressionGraph>, Ptr<data::CorpusBatch>) = 0;

  template <typename T>
  T opt(const std::string& key) {
    return options_->get<T>(key);
  }
  virtual void clear() = 0;
};

class DecoderBase {
protected:
  Ptr<Options> options_;
  std::string prefix_{"decoder"};
  bool inference_{false};
  size_t batchIndex_{1};

public:
  DecoderBase(Ptr<Options> options)
      : options_(options),
        prefix_(options->get<std::string>("prefix", "decoder")),
        inference_(options->get<bool>("inference", false)),
        batchIndex_(options->get<size_t>("index", 1)) {}
  virtual ~DecoderBase() {}

  virtual Ptr<DecoderState> startState(Ptr<ExpressionGraph>,
                                       Ptr<data::CorpusBatch> batch,
                                       std::vector<Ptr<EncoderState>>&)
      = 0;
  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph>, Ptr<DecoderState>) = 0;

  virtual std::tuple<Expr, Expr> groundTruth(Ptr<DecoderState> state,
                                             Ptr<ExpressionGraph> graph,
                                             Ptr<data::CorpusBatch> batch) {
    using namespace keywords;
    size_t bi = batchIndex_;
    int dimVoc = opt<std::vector<int>>("dim-vocabs")[bi];
    int dimEmb = opt<int>("dim-emb");

    auto yEmbFactory = embedding(graph)("dimVocab", dimVoc)("dimEmb", dimEmb);
    if(opt<bool>("tied-embeddings-src") || opt<bool>("tied-embeddings-all"))
      yEmbFactory("prefix", "Wemb");
    else
      yEmbFactory("prefix", prefix_ + "_Wemb");
    if(options_->has("embedding-fix-trg"))
      yEmbFactory("fixed", opt<bool>("embedding-fix-trg"));
    auto yEmb = yEmbFactory.construct();

    auto subBatch = (*batch)[bi];
    int dimBatch = (int)subBatch->batchSize();
    int dimWords = (int)subBatch->batchWidth();

    auto idxFill = [bi](const data::CorpusBatch& b, int* dst) {
      auto& idx = b[bi]->indices();
      for(size_t i = 0; i < idx.size(); ++i)
        dst[i] = (int)idx[i];
    };
    auto chosenEmbeddings = rows(yEmb, subBatch->indices(), idxFill, batch);
    auto y = reshape(chosenEmbeddings, {dimWords, dimBatch, dimEmb});

    auto yMask = graph->batchConstant(
        {dimWords, dimBatch, 1},
        [bi](const data::CorpusBatch& b, float* dst) {
          auto& m = b[bi]->mask();
          std::copy(m.begin(), m.end(), dst);
        },
        batch);
    // labels as a FLOAT tensor, as the reference passes them to cross_entropy
    auto yIdx = graph->batchConstant(
        {(int)subBatch->indices().size(), 1},
        [bi](const data::CorpusBatch& b, float* dst) {
          auto& idx = b[bi]->indices();
          for(size_t i = 0; i < idx.size(); ++i)
            dst[i] = (float)idx[i];
        },
        batch);

    auto yShifted = shift(y, {1, 0, 0});

    state->setTargetEmbeddings(yShifted);
    state->setTargetMask(yMask);
    return std::make_tuple(yMask, yIdx);
  }

  // Decoding: the embeddings of the words chosen in the previous beam-search step become the decoder input,
  // [beam, 1, batch, dimEmb]; no words yet (first step) = zeros [1, 1, batch, dimEmb].   reference: encdec.h:123-153
  virtual void selectEmbeddings(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state, const std::vector<size_t>& words, int dimBatch, int dimBeam) {
    using namespace keywords;
    const int dimEmb = opt<int>("dim-emb");
    auto table = embedding(graph)("dimVocab", opt<std::vector<int>>("dim-vocabs")[batchIndex_])("dimEmb", dimEmb);
    const bool shared = opt<bool>("tied-embeddings-src") || opt<bool>("tied-embeddings-all");
    table("prefix", shared ? std::string("Wemb") : prefix_ + "_Wemb");
    Expr E = table.construct();
    Expr chosen = words.empty() ? graph->constant({1, 1, dimBatch, dimEmb}, init = inits::zeros)
                                : reshape(rows(E, words), {dimBeam, 1, dimBatch, dimEmb});
    state->setTargetEmbeddings(chosen);
  }

  virtual const std::vector<Expr> getAlignments(int i = 0) { return {}; }

  template <typename T>
  T opt(const std::string& key) {
    return options_->get<T>(key);
  }
  virtual void clear() = 0;
};

class EncoderDecoder : public models::ModelBase {
protected:
  Ptr<Options> options_;
  std::string prefix_;
  std::vector<Ptr<EncoderBase>> encoders_;
  std::vector<Ptr<DecoderBase>> decoders_;
  bool inference_{false};

public:
  EncoderDecoder(Ptr<Options> options)
      : options_(options), prefix_(options->get<std::string>("prefix", "")), inference_(options->get<bool>("inference", false)) {}

  std::vector<Ptr<EncoderBase>>& getEncoders() { return encoders_; }
  std::vector<Ptr<DecoderBase>>& getDecoders() { return decoders_; }
  void push_back(Ptr<EncoderBase> encoder) { encoders_.push_back(encoder); }
  void push_back(Ptr<DecoderBase> decoder) { decoders_.push_back(decoder); }

  virtual void clear(Ptr<ExpressionGraph> graph) {
    graph->clear();
    for(auto& enc : encoders_)
      enc->clear();
    for(auto& dec : decoders_)
      dec->clear();
  }

  virtual Ptr<DecoderState> startState(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch) {
    std::vector<Ptr<EncoderState>> encoderStates;
    for(auto& encoder : encoders_)
      encoderStates.push_back(encoder->build(graph, batch));
    return decoders_[0]->startState(graph, batch, encoderStates);
  }

  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state) {
    return decoders_[0]->step(graph, state);
  }

  // One beam-search step: carry over the states of the surviving hypotheses (`hypIndices`, beam-major rows of the
  // previous step), feed the words they chose, run the decoder for ONE position, return log-probabilities.
  // reference: encdec.h:316-335
  virtual Ptr<DecoderState> step(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state, const std::vector<size_t>& hypIndices,
                                 const std::vector<size_t>& embIndices, int dimBatch, int beamSize, bool normalized = true) {
    auto chosen = hypIndices.empty() ? state : state->select(hypIndices, beamSize);
    selectEmbeddings(graph, chosen, embIndices, dimBatch, beamSize);
    chosen->setSingleStep(true);
    auto next = step(graph, chosen);
    if(normalized)  // the fused top-N kernel normalises the rows itself and wants the raw logits
      next->setProbs(logsoftmax(next->getProbs()));
    return next;
  }

  virtual void selectEmbeddings(Ptr<ExpressionGraph> graph, Ptr<DecoderState> state, const std::vector<size_t>& words, int dimBatch, int beamSize) {
    decoders_[0]->selectEmbeddings(graph, state, words, dimBatch, beamSize);
  }

  // Also exposes the logits node of the last build (parity checks compare logits).
  Expr lastLogits() { return lastLogits_; }

  virtual Expr build(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch, bool clearGraph = true) {
    using namespace keywords;
    if(clearGraph)
      clear(graph);

    auto state = startState(graph, batch);

    Expr trgMask, trgIdx;
    std::tie(trgMask, trgIdx) = decoders_[0]->groundTruth(state, graph, batch);

    auto nextState = step(graph, state);
    lastLogits_ = nextState->getProbs();

    std::string costType = opt<std::string>("cost-type");
    float ls = inference_ ? 0.f : opt<float>("label-smoothing");

    return Cost(nextState->getProbs(), trgIdx, trgMask, costType, ls);
  }

  template <typename T>
  T opt(const std::string& key) {
    return options_->get<T>(key);
  }

private:
  Expr lastLogits_;
};

}  // namespace marian
