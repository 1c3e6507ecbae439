// Fused recurrent-cell and attention-score graph nodes.
// reference: src/rnn/cells.cu:9-137 (GRUFastNodeOp, LSTMCellNodeOp,
// LSTMOutputNodeOp) and src/rnn/attention.cu:10-66 (AttentionNodeOp).
// Each has ONE backward closure that produces all input gradients and runs
// regardless of which child is trainable (runBackward override there too).
#include "graph/node_operators_binary.h"
#include "kernels/tensor_operators.h"
#include "rnn/rnn.h"

namespace marian {
namespace rnn {

namespace {
// common shape of the three cell nodes: gather child values / (optional) grads
struct FusedCellNodeOp : public NaryNodeOp {
  FusedCellNodeOp(const std::vector<Expr>& nodes) : NaryNodeOp(nodes) {}

  std::vector<Tensor> inputVals() {
    std::vector<Tensor> inputs;
    for(size_t i = 0; i < children_.size(); ++i)
      inputs.push_back(child(i)->val());
    return inputs;
  }
  std::vector<Tensor> inputGrads() {
    std::vector<Tensor> outputs;
    for(auto child : children_)
      outputs.push_back(child->trainable() ? child->grad() : nullptr);
    return outputs;
  }
  // do not check whether child 0 is trainable
  virtual void runBackward(const NodeOps& ops) {
    for(auto&& op : ops)
      op();
  }
};
}  // namespace

struct GRUFastNodeOp : public FusedCellNodeOp {
  bool final_;
  GRUFastNodeOp(const std::vector<Expr>& nodes, bool final) : FusedCellNodeOp(nodes), final_(final) {}

  NodeOps forwardOps() {
    auto inputs = inputVals();
    return {NodeOp(GRUFastForward(val_, inputs, final_))};
  }
  NodeOps backwardOps() {
    auto inputs = inputVals();
    auto outputs = inputGrads();
    return {NodeOp(GRUFastBackward(outputs, inputs, adj_, final_))};
  }
  virtual size_t hash() {
    if(!hash_) {
      hash_ = NaryNodeOp::hash();
      hash_combine(hash_, final_);
    }
    return hash_;
  }
  const std::string type() { return "GRU-ops"; }
};

Expr gruOps(const std::vector<Expr>& nodes, bool final) {
  return Expression<GRUFastNodeOp>(nodes, final);
}

struct LSTMCellNodeOp : public FusedCellNodeOp {
  LSTMCellNodeOp(const std::vector<Expr>& nodes) : FusedCellNodeOp(nodes) {}
  NodeOps forwardOps() {
    auto inputs = inputVals();
    return {NodeOp(LSTMCellForward(val_, inputs))};
  }
  NodeOps backwardOps() {
    auto inputs = inputVals();
    auto outputs = inputGrads();
    return {NodeOp(LSTMCellBackward(outputs, inputs, adj_))};
  }
  const std::string type() { return "LSTM-cell-ops"; }
};

struct LSTMOutputNodeOp : public FusedCellNodeOp {
  LSTMOutputNodeOp(const std::vector<Expr>& nodes) : FusedCellNodeOp(nodes) {}
  NodeOps forwardOps() {
    auto inputs = inputVals();
    return {NodeOp(LSTMOutputForward(val_, inputs))};
  }
  NodeOps backwardOps() {
    auto inputs = inputVals();
    auto outputs = inputGrads();
    return {NodeOp(LSTMOutputBackward(outputs, inputs, adj_))};
  }
  const std::string type() { return "LSTM-output-ops"; }
};

Expr lstmOpsC(const std::vector<Expr>& nodes) {
  return Expression<LSTMCellNodeOp>(nodes);
}
Expr lstmOpsO(const std::vector<Expr>& nodes) {
  return Expression<LSTMOutputNodeOp>(nodes);
}

// score[j] = sum_k va[k] * tanh(context[j,k] + state[b(j),k])
struct AttentionNodeOp : public NaryNodeOp {
  AttentionNodeOp(const std::vector<Expr>& nodes) : NaryNodeOp(nodes, newShape(nodes)) {}

  static Shape newShape(const std::vector<Expr>& nodes) {
    Shape shape = Shape::broadcast({nodes[1], nodes[2]});
    Shape vaShape = nodes[0]->shape();
    ABORT_IF(vaShape[-2] != shape[-1] || vaShape[-1] != 1, "Wrong size");
    shape.set(-1, 1);
    return shape;
  }

  NodeOps forwardOps() { return {NodeOp(Att(val_, child(0)->val(), child(1)->val(), child(2)->val()))}; }
  NodeOps backwardOps() {
    return {NodeOp(AttBack(child(0)->grad(),
                           child(1)->grad(),
                           child(2)->grad(),
                           child(0)->val(),
                           child(1)->val(),
                           child(2)->val(),
                           adj_))};
  }
  virtual void runBackward(const NodeOps& ops) {
    for(auto&& op : ops)
      op();
  }
  const std::string type() { return "Att-ops"; }
};

Expr attOps(Expr va, Expr context, Expr state) {
  std::vector<Expr> nodes{va, context, state};
  int dimBatch = context->shape()[-2];
  int dimWords = context->shape()[-3];
  int dimBeam = 1;
  if(state->shape().size() > 3)
    dimBeam = state->shape()[-4];
  return reshape(Expression<AttentionNodeOp>(nodes), {dimBeam, 1, dimWords, dimBatch});
}

}  // namespace rnn
}  // namespace marian
