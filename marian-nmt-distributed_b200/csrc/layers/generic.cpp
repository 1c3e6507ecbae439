// Training cost.  reference: src/layers/generic.cpp:5-42
#include "layers/generic.h"

namespace marian {

Expr Cost(Expr logits, Expr indices, Expr mask, std::string costType, float smoothing) {
  using namespace keywords;

  auto ce = cross_entropy(logits, indices);

  if(smoothing > 0) {
    // label smoothing: mix in the mean log-probability of the row
    auto ceq = mean(logsoftmax(logits), axis = -1);
    ce = (1 - smoothing) * ce - smoothing * ceq;
  }

  if(mask)
    ce = ce * mask;

  Expr cost;
  if(costType == "ce-mean" || costType == "cross-entropy") {
    cost = mean(sum(ce, axis = -3), axis = -2);
  } else if(costType == "ce-mean-words") {
    cost = sum(sum(ce, axis = -3), axis = -2) / sum(sum(mask, axis = -3), axis = -2);
  } else if(costType == "ce-sum") {
    cost = sum(sum(ce, axis = -3), axis = -2);
  } else if(costType == "perplexity") {
    cost = exp(sum(sum(ce, axis = -3), axis = -2) / sum(sum(mask, axis = -3), axis = -2));
  } else if(costType == "ce-rescore") {
    cost = -sum(ce, axis = -3);
  } else {
    cost = mean(sum(ce, axis = -3), axis = -2);
  }
  return cost;
}

}  // namespace marian
