// Corpus-level translation: text in, text out, on the parameters of a live training graph.
//
// Reference: Translate<BeamSearch>::run (src/translator/translator.h:22-108) - source corpus -> length-sorted mini-batches
// (maxi-batch-sort src) -> beam search per batch -> best (or n-best) hypothesis per line, written back in corpus order
// (src/translator/output_collector.cpp) as words joined by blanks, n-best lines as "id ||| words ||| F0= cost ||| cost".
// The reference spreads batches over devices with a thread pool; one process per GPU here, a rank translates its own
// share.  The model is the training graph's parameters in inference mode (no dropout).
#pragma once

#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "data/corpus.h"
#include "models/model_factory.h"
#include "translator/beam_search.h"

namespace marian {

class Translator {
public:
  // options: beam-size (12), normalize (0), allow-unk (false), n-best (false), beam-fused-nth (true),
  // mini-batch (1), maxi-batch (1), max-length (1000)
  Translator(Ptr<Options> modelOptions, Ptr<Options> options) : options_(options) {
    auto m = modelOptions->clone();
    m->set("inference", true);
    model_ = models::from_options(m);
    config_.beamSize = options->get<size_t>("beam-size", 12);
    config_.normalize = options->get<float>("normalize", 0.f);
    config_.allowUnk = options->get<bool>("allow-unk", false);
    config_.fusedSelection = options->get<bool>("beam-fused-nth", true);
  }

  const BeamSearch::Config& config() const { return config_; }

  // n best translations of every sentence of `batch` (its source side), best first
  std::vector<std::vector<TranslationResult>> translate(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch, size_t nBest = 1) {
    device::setDevice((int)graph->getDevice());
    BeamSearch search(config_, {Scorer{model_, 1.f}});
    auto histories = search.search(graph, batch);
    std::vector<std::vector<TranslationResult>> out;
    for(auto& h : histories)
      out.push_back(h.nBest(nBest));
    return out;
  }

  // source text file -> translations, one line per input line (n-best: several), in corpus order
  size_t translateFile(Ptr<ExpressionGraph> graph, const std::string& srcPath, Ptr<data::Vocab> srcVocab, Ptr<data::Vocab> trgVocab, const std::string& outPath) {
    auto dataOpts = options_->clone();
    dataOpts->set("shuffle", false);
    dataOpts->set("maxi-batch-sort", std::string("src"));
    if(!options_->has("max-length"))
      dataOpts->set("max-length", 1000);
    if(!options_->has("mini-batch"))
      dataOpts->set("mini-batch", 1);
    if(!options_->has("maxi-batch"))
      dataOpts->set("maxi-batch", 1);
    auto corpus = New<data::Corpus>(std::vector<std::string>{srcPath}, std::vector<Ptr<data::Vocab>>{srcVocab}, dataOpts);
    data::BatchGenerator batches(corpus, dataOpts);
    batches.prepare(false);
    const bool nbest = options_->get<bool>("n-best", false);
    std::map<size_t, std::string> lines;
    while(batches) {
      auto batch = batches.next();
      auto results = translate(graph, batch, nbest ? config_.beamSize : 1);
      for(size_t i = 0; i < results.size(); ++i) {
        size_t id = batch->getSentenceIds()[i];
        std::ostringstream line;
        for(size_t r = 0; r < results[i].size(); ++r) {
          std::string text = join((*trgVocab)(results[i][r].words));
          if(nbest)
            line << (r ? "\n" : "") << id << " ||| " << text << " ||| F0= " << results[i][r].rawCost << " ||| " << results[i][r].cost;
          else
            line << text;
        }
        lines[id] = line.str();
      }
    }
    std::ofstream out(outPath);
    ABORT_IF(!out.good(), "Cannot write translations to:", outPath);
    for(auto& kv : lines)
      out << kv.second << "\n";
    return lines.size();
  }

private:
  static std::string join(const std::vector<std::string>& words) {
    std::string s;
    for(size_t i = 0; i < words.size(); ++i)
      s += (i ? " " : "") + words[i];
    return s;
  }

  Ptr<Options> options_;
  Ptr<EncoderDecoder> model_;
  BeamSearch::Config config_;
};

}  // namespace marian
