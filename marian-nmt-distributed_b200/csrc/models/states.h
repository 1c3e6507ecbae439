// Encoder / decoder state carriers.  reference: src/models/states.h:7-81
#pragma once

#include "data/batch.h"
#include "graph/expression_graph.h"
#include "rnn/states.h"

namespace marian {

class EncoderState {
private:
  Expr context_;
  Expr mask_;
  Ptr<data::CorpusBatch> batch_;

public:
  EncoderState(Expr context, Expr mask, Ptr<data::CorpusBatch> batch) : context_(context), mask_(mask), batch_(batch) {}
  virtual ~EncoderState() {}

  virtual Expr getContext() { return context_; }
  virtual Expr getAttended() { return context_; }
  virtual Expr getMask() { return mask_; }
  virtual const std::vector<size_t>& getSourceWords() { return batch_->front()->indices(); }
};

class DecoderState {
protected:
  rnn::States states_;
  Expr probs_;
  std::vector<Ptr<EncoderState>> encStates_;
  Expr targetEmbeddings_;
  Expr targetMask_;
  bool singleStep_{false};

public:
  DecoderState(const rnn::States& states, Expr probs, std::vector<Ptr<EncoderState>>& encStates)
      : states_(states), probs_(probs), encStates_(encStates) {}
  virtual ~DecoderState() {}

  virtual std::vector<Ptr<EncoderState>>& getEncoderStates() { return encStates_; }

  virtual Expr getProbs() { return probs_; }
  virtual void setProbs(Expr probs) { probs_ = probs; }

  virtual const rnn::States& getStates() { return states_; }

  // Beam search: the state of the hypotheses `selIdx` (rows of the [beam, batch] layout, beam-major) re-packed as
  // [beamSize, .., batch, ..].  reference: states.h:52-54
  virtual Ptr<DecoderState> select(const std::vector<size_t>& selIdx, int beamSize) {
    return New<DecoderState>(states_.select(selIdx, beamSize), probs_, encStates_);
  }
  // hook for decoders that forbid words at some positions (reference: states.h:75); none of the models here does
  virtual void blacklist(Expr /*totalCosts*/, Ptr<data::CorpusBatch> /*batch*/) {}

  virtual Expr getTargetEmbeddings() { return targetEmbeddings_; }
  virtual void setTargetEmbeddings(Expr targetEmbeddings) { targetEmbeddings_ = targetEmbeddings; }

  virtual Expr getTargetMask() { return targetMask_; }
  virtual void setTargetMask(Expr targetMask) { targetMask_ = targetMask; }

  virtual bool doSingleStep() { return singleStep_; }
  virtual void setSingleStep(bool singleStep = true) { singleStep_ = singleStep; }
};

}  // namespace marian
