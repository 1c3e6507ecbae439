// Validation on a held-out corpus: forward passes only, on the training graph's parameters.
//
// Reference: CrossEntropyValidator (src/training/validator.h:108-176): an inference-mode model ("inference" = true:
// no dropout; "cost-type" = ce-sum) is built for every mini-batch of the validation set on the TRAINING graph,
// forward() runs, the batch costs are summed and reported as
//   cross-entropy / ce-mean : cost / sentences      ce-mean-words : cost / target words
//   perplexity              : exp(cost / target words)            ce-sum : cost.
// The reference spreads the batches over the training graphs of all devices with a thread pool; one process per
// GPU here, so a rank validates on its own graph (validation sets are small next to training steps).
// Batches come from data/corpus.h (length-sorted, not shuffled).  The CUDA-graph plans of training steps stay
// valid: they re-create every workspace tensor they read.
#pragma once

#include <cmath>
#include <string>
#include <vector>

#include "data/corpus.h"
#include "graph/expression_graph.h"
#include "models/model_factory.h"

namespace marian {

class CrossEntropyValidator {
public:
  CrossEntropyValidator(std::vector<Ptr<data::Vocab>> vocabs, Ptr<Options> options) : vocabs_(vocabs), options_(options) {
    auto opts = options->clone();
    opts->set("inference", true);
    opts->set("cost-type", "ce-sum");
    builder_ = models::from_options(opts);
  }

  std::string type() const { return options_->get<std::string>("cost-type", "ce-mean"); }

  // `paths`: one text file per side.  Returns the metric selected by "cost-type"; fills the totals if asked.
  float validate(Ptr<ExpressionGraph> graph, const std::vector<std::string>& paths, float* costSum = nullptr, size_t* sentences = nullptr, size_t* targetWords = nullptr) {
    auto dataOpts = options_->clone();
    dataOpts->set("shuffle", false);
    if(options_->has("valid-mini-batch"))
      dataOpts->set("mini-batch", options_->get<int>("valid-mini-batch"));
    if(options_->has("valid-max-length"))
      dataOpts->set("max-length", options_->get<int>("valid-max-length"));
    auto corpus = New<data::Corpus>(paths, vocabs_, dataOpts);
    data::BatchGenerator batches(corpus, dataOpts);
    batches.prepare(false);

    device::setDevice((int)graph->getDevice());
    graph->setBackwardSplit(nullptr, nullptr);
    double cost = 0;
    size_t samples = 0, words = 0;
    while(batches) {
      auto batch = batches.next();
      auto costNode = builder_->build(graph, batch);
      graph->forward();
      cost += costNode->scalar();  // blocking read-back, as the reference
      samples += batch->size();
      words += batch->back()->batchWords();
    }
    if(costSum)
      *costSum = (float)cost;
    if(sentences)
      *sentences = samples;
    if(targetWords)
      *targetWords = words;
    auto ctype = type();
    if(ctype == "perplexity")
      return (float)std::exp(cost / (double)words);
    if(ctype == "ce-mean-words")
      return (float)(cost / (double)words);
    if(ctype == "ce-sum")
      return (float)cost;
    return (float)(cost / (double)samples);
  }

private:
  std::vector<Ptr<data::Vocab>> vocabs_;
  Ptr<Options> options_;
  Ptr<EncoderDecoder> builder_;
};

}  // namespace marian
