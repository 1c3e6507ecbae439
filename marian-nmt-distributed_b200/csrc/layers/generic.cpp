// Training cost on top of the logits.
// Semantics of the reference's Cost() (src/layers/generic.cpp:5-42): per-token cross entropy,
// optional label smoothing against the uniform distribution, masked, then reduced over time
// (axis -3) and batch (axis -2) according to --cost-type.
#include "layers/generic.h"

namespace marian {

Expr Cost(Expr logits, Expr indices, Expr mask, std::string costType, float smoothing) {
  using namespace keywords;

  Expr tokenLoss = cross_entropy(logits, indices);
  if(smoothing > 0) {
    // smoothed target = (1 - s) * one-hot + s * uniform; the uniform part costs the mean log-prob
    Expr uniformPart = mean(logsoftmax(logits), axis = -1);
    tokenLoss = (1 - smoothing) * tokenLoss - smoothing * uniformPart;
  }
  if(mask)
    tokenLoss = tokenLoss * mask;

  // [1, 1, B, 1] loss per sentence; the remaining reductions are over the batch axis
  Expr perSentence = sum(tokenLoss, axis = -3);
  auto overBatch = [](Expr x) { return sum(x, keywords::axis = -2); };
  auto tokenCount = [&]() { return overBatch(sum(mask, axis = -3)); };

  if(costType == "ce-sum")
    return overBatch(perSentence);
  if(costType == "ce-mean-words")
    return overBatch(perSentence) / tokenCount();
  if(costType == "perplexity")
    return exp(overBatch(perSentence) / tokenCount());
  if(costType == "ce-rescore")
    return -perSentence;
  // "ce-mean", "cross-entropy" and anything unknown: mean sentence loss
  return mean(perSentence, axis = -2);
}

}  // namespace marian
