// Beam-search decoding on the training graph's parameters (validation / decoding).
//
// Behavioural contract = the reference's translator core:
//   src/translator/hypothesis.h:9-42    a hypothesis = (previous hypothesis, word, row of the previous step's state, cost)
//   src/translator/history.h:11-78      per sentence: the beams of every step + a queue of finished hypotheses ranked by
//                                       cost / length^normalize; n-best by back-tracking
//   src/translator/scorers.h:9-100      a scorer wraps an encoder-decoder: startState / step -> log-probabilities
//   src/translator/beam_search.h:28-225 the search loop, including its conventions: all sentences of a batch advance together,
//                                       state rows are beam-major ([beam, batch]), the scored tensor is regrouped per sentence
//                                       before the n-best selection, keys = (sentence*beam + hypothesis)*V + word, finished
//                                       hypotheses (word 0 = </s>) leave the beam, the beam width of the next step is the
//                                       widest surviving beam, a sentence is cut after 3 x source length steps.
// Organisation is this repo's own: hypotheses live in one flat arena (indices instead of shared pointers), the loop is
// split into "build the step's nodes" / "select" / "book-keeping", and the selection has two implementations behind one
// flag: the reference's node sequence (logsoftmax + add + transpose, then NthElementRanges) and the fused operator
// NthElementLogSoftmax that reads the raw logits once (kernels/nth_element.cu).  Both are run against each other and
// against the CPU oracle in tests/test_translator.py / tests/test_gpu_translator.py.
#pragma once

#include <algorithm>
#include <cmath>
#include <limits>
#include <queue>
#include <vector>

#include "data/batch.h"
#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "kernels/tensor_operators.h"
#include "models/encdec.h"

namespace marian {

constexpr size_t kEosId = 0;  // reference: src/data/types.h (EOS_ID 0, UNK_ID 1)
constexpr size_t kUnkId = 1;

struct Hypothesis {
  int prev;            // arena index of the hypothesis this one extends, -1 for the empty start hypothesis
  size_t word;         // word chosen in this step
  size_t prevStateRow; // row ([beam, batch], beam-major) of the decoder state this hypothesis continues
  float cost;          // accumulated log-probability
};

typedef std::vector<int> Beam;  // arena indices, best first
typedef std::vector<Beam> Beams;

struct TranslationResult {
  std::vector<size_t> words;  // without the start symbol, including </s> if the hypothesis ended
  float cost;                 // accumulated log-probability divided by length^normalize
  float rawCost;              // accumulated log-probability
};

class History {
public:
  History(const std::vector<Hypothesis>* arena, float alpha) : arena_(arena), alpha_(alpha) {}

  float lengthPenalty(size_t length) const { return std::pow((float)length, alpha_); }

  // Records the beam of one step; hypotheses that just produced </s> (or all of them when `last`) become candidates
  // for the final ranking, scored by cost / (number of steps so far)^alpha.
  void add(const Beam& beam, bool last = false) {
    if(!beam.empty() && (*arena_)[beam.back()].prev >= 0) {
      for(size_t j = 0; j < beam.size(); ++j) {
        const Hypothesis& h = (*arena_)[beam[j]];
        if(h.word == kEosId || last)
          finished_.push(Finished{h.cost / lengthPenalty(steps_.size()), steps_.size(), j});
      }
    }
    steps_.push_back(beam);
  }

  size_t size() const { return steps_.size(); }

  std::vector<TranslationResult> nBest(size_t n) const {
    std::vector<TranslationResult> out;
    auto queue = finished_;
    while(out.size() < n && !queue.empty()) {
      Finished f = queue.top();
      queue.pop();
      TranslationResult r;
      int at = steps_[f.step][f.slot];
      r.rawCost = (*arena_)[at].cost;
      r.cost = f.score;
      for(; (*arena_)[at].prev >= 0; at = (*arena_)[at].prev)
        r.words.push_back((*arena_)[at].word);
      std::reverse(r.words.begin(), r.words.end());
      out.push_back(r);
    }
    return out;
  }

private:
  struct Finished {
    float score;
    size_t step, slot;
    bool operator<(const Finished& o) const { return score < o.score; }
  };
  const std::vector<Hypothesis>* arena_;
  float alpha_;
  std::vector<Beam> steps_;
  std::priority_queue<Finished> finished_;
};

// One model scored with weight 1 is the common case; several (an ensemble over one vocabulary) add up their weighted
// log-probabilities as in the reference's scorer list.
struct Scorer {
  Ptr<EncoderDecoder> model;
  float weight{1.f};
};

class BeamSearch {
public:
  struct Config {
    size_t beamSize{12};
    float normalize{0.f};
    bool allowUnk{false};
    bool fusedSelection{true};  // NthElementLogSoftmax on the raw logits (single scorer, weight 1)
    int maxLengthFactor{3};
  };

  BeamSearch(const Config& config, const std::vector<Scorer>& scorers) : config_(config), scorers_(scorers) {
    ABORT_IF(scorers_.empty(), "beam search needs at least one scorer");
    ABORT_IF(config_.beamSize < 1, "beam-size must be positive");
  }

  // Arena of the last search (the histories returned by search() point into it).
  const std::vector<Hypothesis>& hypotheses() const { return arena_; }

  std::vector<History> search(Ptr<ExpressionGraph> graph, Ptr<data::CorpusBatch> batch) {
    const int sentences = (int)batch->size();
    const size_t maxSteps = (size_t)config_.maxLengthFactor * batch->front()->batchWidth();
    const bool fused = config_.fusedSelection && scorers_.size() == 1 && scorers_[0].weight == 1.f;

    arena_.clear();
    arena_.push_back(Hypothesis{-1, 0, 0, 0.f});  // the shared empty start hypothesis
    std::vector<History> histories(sentences, History(&arena_, config_.normalize));
    size_t width = config_.beamSize;
    Beams beams(sentences, Beam(width, 0));
    for(int s = 0; s < sentences; ++s)
      histories[s].add(beams[s]);

    const bool wasInference = graph->inference();
    graph->setInference(true);  // nodes of finished steps are released as soon as nothing refers to them
    graph->setBackwardSplit(nullptr, nullptr);
    {
      auto gemm = graph->getBackend()->getGemmHandle();
      gemmAllowShadowOnly(gemm, false);  // every fp32 value stays readable (the selection reads the logits)
      gemmPrepareStep(gemm);             // bf16 copy of the parameters up to date (e.g. right after a checkpoint load)
    }
    for(auto& sc : scorers_)
      sc.model->clear(graph);
    std::vector<Ptr<DecoderState>> states;
    for(auto& sc : scorers_)
      states.push_back(sc.model->startState(graph, batch));

    Expr unkPenalty;  // [1, V]: lowest() at <unk>, 0 elsewhere (x + lowest() == lowest() in float arithmetic)
    bool first = true, cut = false;
    do {
      // ---- rows of the previous step that survive, the words they chose, their costs (beam-major) ----
      std::vector<size_t> stateRows, words;
      std::vector<float> costs;
      if(first) {
        costs.assign(sentences, 0.f);
      } else {
        for(size_t i = 0; i < width; ++i)
          for(int s = 0; s < sentences; ++s) {
            if(i < beams[s].size()) {
              const Hypothesis& h = arena_[beams[s][i]];
              stateRows.push_back(h.prevStateRow);
              words.push_back(h.word);
              costs.push_back(h.cost);
            } else {  // padding row of a sentence with fewer live hypotheses
              stateRows.push_back(0);
              words.push_back(0);
              costs.push_back(-9999.f);
            }
          }
      }

      // ---- decoder step for every scorer ----
      for(size_t i = 0; i < scorers_.size(); ++i)
        states[i] = scorers_[i].model->step(graph, states[i], stateRows, words, sentences, (int)width, /*normalized=*/!fused);

      std::vector<float> outCosts;
      std::vector<unsigned> outKeys;
      const int V = states[0]->getProbs()->shape()[-1];
      if(fused) {
        Expr logits = states[0]->getProbs();
        runForward(graph, first);
        NthElementLogSoftmax(logits->val(), costs, sentences, (int)width, (int)width, first, config_.allowUnk ? -1 : (int)kUnkId, outCosts, outKeys);
      } else {
        using namespace keywords;
        Expr total = first ? graph->constant({1, 1, 1, 1}, init = inits::from_value(0))
                           : graph->constant({(int)width, 1, sentences, 1}, init = inits::from_vector(costs));
        for(size_t i = 0; i < scorers_.size(); ++i)
          total = scorers_[i].weight != 1.f ? total + scorers_[i].weight * states[i]->getProbs() : total + states[i]->getProbs();
        if(!config_.allowUnk) {
          if(!unkPenalty) {
            std::vector<float> pen(V, 0.f);
            pen[kUnkId] = std::numeric_limits<float>::lowest();
            unkPenalty = graph->constant({1, V}, init = inits::from_vector(pen));
          }
          total = total + unkPenalty;
        }
        if(sentences > 1 && width > 1 && !first)
          total = transpose(total, {2, 1, 0, 3});  // [batch, 1, beam, V]: a sentence's hypotheses become contiguous
        runForward(graph, first);
        const int rowsPerSentence = first ? 1 : (int)width;
        std::vector<int> rangeFirst(sentences + 1), cumN(sentences + 1);
        for(int s = 0; s <= sentences; ++s) {
          rangeFirst[s] = s * rowsPerSentence * V;
          cumN[s] = s * (int)width;
        }
        NthElementRanges(total->val(), rangeFirst, cumN, outCosts, outKeys);
      }

      // ---- keys -> hypotheses (reference toHyps, beam_search.h:28-76) ----
      Beams next(sentences);
      for(size_t i = 0; i < outKeys.size(); ++i) {
        const int s = (int)(i / width);
        if(next[s].size() >= beams[s].size())
          continue;  // finished sentences keep nothing; narrower beams keep their best
        const size_t word = outKeys[i] % (unsigned)V;
        const size_t row = outKeys[i] / (unsigned)V;  // sentence-major row: sentence*width + hypothesis
        size_t slot = row % width;
        const size_t stateRow = first ? row : (row / width) + slot * (size_t)sentences;  // back to beam-major
        if(slot >= beams[s].size())
          slot = slot % beams[s].size();
        if(first)
          slot = 0;
        arena_.push_back(Hypothesis{beams[s][slot], word, stateRow, outCosts[i]});
        next[s].push_back((int)arena_.size() - 1);
      }

      // ---- finished hypotheses leave the beam ----
      Beams live(sentences);
      for(int s = 0; s < sentences; ++s)
        for(int h : next[s])
          if(arena_[h].word != kEosId)
            live[s].push_back(h);
      for(int s = 0; s < sentences; ++s)
        if(!next[s].empty()) {
          cut = cut || histories[s].size() >= maxSteps;
          histories[s].add(next[s], live[s].empty() || cut);
        }
      beams = live;

      if(!first) {
        width = 0;
        for(auto& b : beams)
          width = std::max(width, b.size());
      }
      first = false;
    } while(width != 0 && !cut);

    states.clear();
    unkPenalty = nullptr;
    for(auto& sc : scorers_)
      sc.model->clear(graph);
    graph->setInference(wasInference);
    return histories;
  }

private:
  // The reference runs forward() for the first step and forwardNext() afterwards; here every step is a full forward():
  // it also rewinds the per-pass operand scratch of the products (bf16 shadows / packed operands), which a decoding run
  // of a hundred steps would otherwise keep growing - nothing of an earlier step's scratch is read again, the decoder
  // states are re-gathered per step.
  static void runForward(Ptr<ExpressionGraph> graph, bool first) {
    graph->forward();
    if(first)  // parameters loaded from a checkpoint and never stepped: the first forward() announced their range
      gemmPrepareStep(graph->getBackend()->getGemmHandle());
  }

  Config config_;
  std::vector<Scorer> scorers_;
  std::vector<Hypothesis> arena_;
};

}  // namespace marian
