// Recurrent layers: cells (tanh / GRU / LSTM), Bahdanau attention as a cell
// input, stacked ("deep transition") cells, the time loop, and their builders.
//
// Structure, parameter names/creation order and formulas follow the reference:
//   src/rnn/types.h:100-266     Stackable / CellInput / Cell / StackedCell
//   src/rnn/cells.h:19-609      Tanh, GRU, FastLSTM (fused gate kernels via gruOps/lstmOps)
//   src/rnn/attention.h:16-168  GlobalAttention (attOps + masked softmax + scalar_product)
//   src/rnn/rnn.h:54-255        SingleLayerRNN time loop, RNN layer stack
//   src/rnn/constructors.h      builders (re-implemented without the YAML Accumulator)
#pragma once

#include <algorithm>
#include <functional>

#include "common/options.h"
#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "layers/generic.h"
#include "models/states.h"
#include "rnn/states.h"

namespace marian {
namespace rnn {

enum struct dir : int { forward, backward, alternating_forward, alternating_backward };

// fused nodes (rnn/cells.cpp)
Expr gruOps(const std::vector<Expr>& nodes, bool final = false);
Expr lstmOpsC(const std::vector<Expr>& nodes);
Expr lstmOpsO(const std::vector<Expr>& nodes);
Expr attOps(Expr va, Expr context, Expr state);

class RNN;

class Stackable : public std::enable_shared_from_this<Stackable> {
protected:
  Ptr<Options> options_;

public:
  Stackable(Ptr<Options> options) : options_(options) {}
  virtual ~Stackable() {}

  template <typename Cast>
  Ptr<Cast> as() {
    return std::dynamic_pointer_cast<Cast>(shared_from_this());
  }
  template <typename Cast>
  bool is() {
    return as<Cast>() != nullptr;
  }
  Ptr<Options> getOptions() { return options_; }
  template <typename T>
  T opt(const std::string& key) {
    return options_->get<T>(key);
  }
  template <typename T>
  T opt(const std::string& key, T defaultValue) {
    return options_->get<T>(key, defaultValue);
  }
  virtual void clear() = 0;
};

class CellInput : public Stackable {
public:
  CellInput(Ptr<Options> options) : Stackable(options) {}
  virtual Expr apply(State) = 0;
  virtual int dimOutput() = 0;
};

class Cell : public Stackable {
protected:
  std::vector<std::function<Expr(Ptr<rnn::RNN>)>> lazyInputs_;

public:
  Cell(Ptr<Options> options) : Stackable(options) {}

  State apply(std::vector<Expr> inputs, State state, Expr mask = nullptr) {
    return applyState(applyInput(inputs), state, mask);
  }
  virtual std::vector<Expr> getLazyInputs(Ptr<rnn::RNN> parent) {
    std::vector<Expr> inputs;
    for(auto lazy : lazyInputs_)
      inputs.push_back(lazy(parent));
    return inputs;
  }
  virtual void setLazyInputs(std::vector<std::function<Expr(Ptr<rnn::RNN>)>> lazy) { lazyInputs_ = lazy; }

  virtual std::vector<Expr> applyInput(std::vector<Expr> inputs) = 0;
  virtual State applyState(std::vector<Expr>, State, Expr = nullptr) = 0;
  virtual void clear() {}
};

// Deep-transition cell: cell 0 consumes the layer input, later cells consume
// what the CellInputs (attention) in between produced.  reference: types.h:201-264
class StackedCell : public Cell {
protected:
  std::vector<Ptr<Stackable>> stackables_;
  std::vector<Expr> lastInputs_;

public:
  StackedCell(Ptr<ExpressionGraph>, Ptr<Options> options) : Cell(options) {}

  void push_back(Ptr<Stackable> stackable) { stackables_.push_back(stackable); }

  virtual std::vector<Expr> applyInput(std::vector<Expr> inputs) {
    return stackables_[0]->as<Cell>()->applyInput(inputs);
  }

  virtual State applyState(std::vector<Expr> mappedInputs, State state, Expr mask = nullptr) {
    State hidden = stackables_[0]->as<Cell>()->applyState(mappedInputs, state, mask);
    for(size_t i = 1; i < stackables_.size(); ++i) {
      if(stackables_[i]->is<Cell>()) {
        auto hiddenNext = stackables_[i]->as<Cell>()->apply(lastInputs_, hidden, mask);
        lastInputs_.clear();
        hidden = hiddenNext;
      } else {
        lastInputs_.push_back(stackables_[i]->as<CellInput>()->apply(hidden));
      }
    }
    return hidden;
  }

  Ptr<Stackable> operator[](int i) { return stackables_[i]; }
  Ptr<Stackable> at(int i) { return stackables_[i]; }

  virtual void clear() {
    for(auto s : stackables_)
      s->clear();
  }
  virtual std::vector<Expr> getLazyInputs(Ptr<rnn::RNN> parent) {
    return stackables_[0]->as<Cell>()->getLazyInputs(parent);
  }
  virtual void setLazyInputs(std::vector<std::function<Expr(Ptr<rnn::RNN>)>> lazy) {
    stackables_[0]->as<Cell>()->setLazyInputs(lazy);
  }
};

// ---------------------------------------------------------------------------
// cells
// ---------------------------------------------------------------------------

namespace detail {
inline Expr joinInputs(const std::vector<Expr>& inputs) {
  if(inputs.size() > 1)
    return concatenate(inputs, keywords::axis = -1);
  return inputs.front();
}
}  // namespace detail

// h' = tanh(x W + h U + b)          reference: cells.h:19-116
class Tanh : public Cell {
private:
  Expr U_, W_, b_;
  Expr gamma1_, gamma2_;
  bool layerNorm_;
  float dropout_;
  Expr dropMaskX_, dropMaskS_;

public:
  Tanh(Ptr<ExpressionGraph> graph, Ptr<Options> options) : Cell(options) {
    int dimInput = opt<int>("dimInput");
    int dimState = opt<int>("dimState");
    std::string prefix = opt<std::string>("prefix");
    layerNorm_ = opt<bool>("layer-normalization", false);
    dropout_ = opt<float>("dropout", 0);

    U_ = graph->param(prefix + "_U", {dimState, dimState}, keywords::init = inits::glorot_uniform);
    if(dimInput)
      W_ = graph->param(prefix + "_W", {dimInput, dimState}, keywords::init = inits::glorot_uniform);
    b_ = graph->param(prefix + "_b", {1, dimState}, keywords::init = inits::zeros);

    if(dropout_ > 0.0f) {
      if(dimInput)
        dropMaskX_ = graph->dropout(dropout_, {1, dimInput});
      dropMaskS_ = graph->dropout(dropout_, {1, dimState});
    }
    if(layerNorm_) {
      if(dimInput)
        gamma1_ = graph->param(prefix + "_gamma1", {1, 3 * dimState}, keywords::init = inits::from_value(1.f));
      gamma2_ = graph->param(prefix + "_gamma2", {1, 3 * dimState}, keywords::init = inits::from_value(1.f));
    }
  }

  std::vector<Expr> applyInput(std::vector<Expr> inputs) {
    if(inputs.empty())
      return {};
    Expr input = detail::joinInputs(inputs);
    if(dropMaskX_)
      input = dropout(input, keywords::mask = dropMaskX_);
    auto xW = dot(input, W_);
    if(layerNorm_)
      xW = layer_norm(xW, gamma1_);
    return {xW};
  }

  State applyState(std::vector<Expr> xWs, State state, Expr mask = nullptr) {
    Expr recState = state.output;
    auto stateDropped = recState;
    if(dropMaskS_)
      stateDropped = dropout(recState, keywords::mask = dropMaskS_);
    auto sU = dot(stateDropped, U_);
    if(layerNorm_)
      sU = layer_norm(sU, gamma2_);

    Expr output = xWs.empty() ? tanh(sU, b_) : tanh(xWs.front(), sU, b_);
    if(mask)
      return {output * mask, nullptr};
    return {output, state.cell};
  }
};

// Gated recurrent unit with the fused gate kernel.   reference: cells.h:120-256
// U|Ux, W|Wx, b|bx are concatenated once per tape into [.., 3*dimState].
class GRU : public Cell {
protected:
  Expr U_, W_, b_;
  Expr gamma1_, gamma2_;
  bool final_;
  bool layerNorm_;
  float dropout_;
  Expr dropMaskX_, dropMaskS_;
  Expr fakeInput_;

public:
  GRU(Ptr<ExpressionGraph> graph, Ptr<Options> options) : Cell(options) {
    int dimInput = opt<int>("dimInput");
    int dimState = opt<int>("dimState");
    std::string prefix = opt<std::string>("prefix");
    layerNorm_ = opt<bool>("layer-normalization", false);
    dropout_ = opt<float>("dropout", 0);
    final_ = opt<bool>("final", false);

    auto U = graph->param(prefix + "_U", {dimState, 2 * dimState}, keywords::init = inits::glorot_uniform);
    auto Ux = graph->param(prefix + "_Ux", {dimState, dimState}, keywords::init = inits::glorot_uniform);
    U_ = concatenate({U, Ux}, keywords::axis = -1);

    if(dimInput > 0) {
      auto W = graph->param(prefix + "_W", {dimInput, 2 * dimState}, keywords::init = inits::glorot_uniform);
      auto Wx = graph->param(prefix + "_Wx", {dimInput, dimState}, keywords::init = inits::glorot_uniform);
      W_ = concatenate({W, Wx}, keywords::axis = -1);
    }

    auto b = graph->param(prefix + "_b", {1, 2 * dimState}, keywords::init = inits::zeros);
    auto bx = graph->param(prefix + "_bx", {1, dimState}, keywords::init = inits::zeros);
    b_ = concatenate({b, bx}, keywords::axis = -1);

    if(dropout_ > 0.0f) {
      if(dimInput)
        dropMaskX_ = graph->dropout(dropout_, {1, dimInput});
      dropMaskS_ = graph->dropout(dropout_, {1, dimState});
    }
    if(layerNorm_) {
      if(dimInput)
        gamma1_ = graph->param(prefix + "_gamma1", {1, 3 * dimState}, keywords::init = inits::from_value(1.f));
      gamma2_ = graph->param(prefix + "_gamma2", {1, 3 * dimState}, keywords::init = inits::from_value(1.f));
    }
  }

  virtual std::vector<Expr> applyInput(std::vector<Expr> inputs) {
    if(inputs.empty())
      return {};
    Expr input = detail::joinInputs(inputs);
    if(dropMaskX_)
      input = dropout(input, keywords::mask = dropMaskX_);
    auto xW = dot(input, W_);
    if(layerNorm_)
      xW = layer_norm(xW, gamma1_);
    return {xW};
  }

  virtual State applyState(std::vector<Expr> xWs, State state, Expr mask = nullptr) {
    auto stateOrig = state.output;
    auto stateDropped = stateOrig;
    if(dropMaskS_)
      stateDropped = dropout(stateOrig, keywords::mask = dropMaskS_);

    auto sU = dot(stateDropped, U_);
    if(layerNorm_)
      sU = layer_norm(sU, gamma2_);

    Expr xW;
    if(xWs.empty()) {
      // transition cell without input: an all-zero xW of the right shape
      if(!fakeInput_ || fakeInput_->shape() != sU->shape())
        fakeInput_ = sU->graph()->constant(sU->shape(), keywords::init = inits::zeros);
      xW = fakeInput_;
    } else {
      xW = xWs.front();
    }

    auto output = mask ? gruOps({stateOrig, xW, sU, b_, mask}, final_) : gruOps({stateOrig, xW, sU, b_}, final_);
    return {output, state.cell};
  }

  virtual void clear() { fakeInput_ = nullptr; }
};

// LSTM with fused cell/output kernels.   reference: cells.h:493-609
class FastLSTM : public Cell {
protected:
  Expr U_, W_, b_;
  Expr gamma1_, gamma2_;
  bool layerNorm_;
  float dropout_;
  Expr dropMaskX_, dropMaskS_;
  Expr fakeInput_;

public:
  FastLSTM(Ptr<ExpressionGraph> graph, Ptr<Options> options) : Cell(options) {
    int dimInput = opt<int>("dimInput");
    int dimState = opt<int>("dimState");
    std::string prefix = opt<std::string>("prefix");
    layerNorm_ = opt<bool>("layer-normalization", false);
    dropout_ = opt<float>("dropout", 0);

    U_ = graph->param(prefix + "_U", {dimState, 4 * dimState}, keywords::init = inits::glorot_uniform);
    if(dimInput)
      W_ = graph->param(prefix + "_W", {dimInput, 4 * dimState}, keywords::init = inits::glorot_uniform);
    b_ = graph->param(prefix + "_b", {1, 4 * dimState}, keywords::init = inits::zeros);

    if(dropout_ > 0.0f) {
      if(dimInput)
        dropMaskX_ = graph->dropout(dropout_, {1, dimInput});
      dropMaskS_ = graph->dropout(dropout_, {1, dimState});
    }
    if(layerNorm_) {
      if(dimInput)
        gamma1_ = graph->param(prefix + "_gamma1", {1, 4 * dimState}, keywords::init = inits::from_value(1.f));
      gamma2_ = graph->param(prefix + "_gamma2", {1, 4 * dimState}, keywords::init = inits::from_value(1.f));
    }
  }

  virtual std::vector<Expr> applyInput(std::vector<Expr> inputs) {
    if(inputs.empty())
      return {};
    Expr input = detail::joinInputs(inputs);
    if(dropMaskX_)
      input = dropout(input, keywords::mask = dropMaskX_);
    auto xW = dot(input, W_);
    if(layerNorm_)
      xW = layer_norm(xW, gamma1_);
    return {xW};
  }

  virtual State applyState(std::vector<Expr> xWs, State state, Expr mask = nullptr) {
    auto recState = state.output;
    auto cellState = state.cell;

    auto recStateDropped = recState;
    if(dropMaskS_)
      recStateDropped = dropout(recState, keywords::mask = dropMaskS_);

    auto sU = dot(recStateDropped, U_);
    if(layerNorm_)
      sU = layer_norm(sU, gamma2_);

    Expr xW;
    if(xWs.empty()) {
      if(!fakeInput_ || fakeInput_->shape() != sU->shape())
        fakeInput_ = sU->graph()->constant(sU->shape(), keywords::init = inits::zeros);
      xW = fakeInput_;
    } else {
      xW = xWs.front();
    }

    auto nextCellState = mask ? lstmOpsC({cellState, xW, sU, b_, mask}) : lstmOpsC({cellState, xW, sU, b_});
    auto nextRecState = lstmOpsO({nextCellState, xW, sU, b_});
    return {nextRecState, nextCellState};
  }

  virtual void clear() { fakeInput_ = nullptr; }
};
using LSTM = FastLSTM;

// ---------------------------------------------------------------------------
// Bahdanau (MLP) attention over the encoder context.   reference: attention.h:16-168
// ---------------------------------------------------------------------------
class GlobalAttention : public CellInput {
private:
  Expr Wa_, ba_, Ua_, va_;
  Expr gammaContext_, gammaState_;
  Ptr<EncoderState> encState_;
  Expr softmaxMask_;
  Expr mappedContext_;
  std::vector<Expr> contexts_;
  std::vector<Expr> alignments_;
  bool layerNorm_;
  float dropout_;
  Expr contextDropped_;
  Expr dropMaskContext_, dropMaskState_;

public:
  GlobalAttention(Ptr<ExpressionGraph> graph, Ptr<Options> options, Ptr<EncoderState> encState)
      : CellInput(options), encState_(encState), contextDropped_(encState->getContext()) {
    int dimDecState = opt<int>("dimState");
    dropout_ = opt<float>("dropout", 0);
    layerNorm_ = opt<bool>("layer-normalization", false);
    std::string prefix = opt<std::string>("prefix");

    int dimEncState = encState_->getContext()->shape()[-1];

    Wa_ = graph->param(prefix + "_W_comb_att", {dimDecState, dimEncState}, keywords::init = inits::glorot_uniform);
    Ua_ = graph->param(prefix + "_Wc_att", {dimEncState, dimEncState}, keywords::init = inits::glorot_uniform);
    va_ = graph->param(prefix + "_U_att", {dimEncState, 1}, keywords::init = inits::glorot_uniform);
    ba_ = graph->param(prefix + "_b_att", {1, dimEncState}, keywords::init = inits::zeros);

    if(dropout_ > 0.0f) {
      dropMaskContext_ = graph->dropout(dropout_, {1, dimEncState});
      dropMaskState_ = graph->dropout(dropout_, {1, dimDecState});
    }
    if(dropMaskContext_)
      contextDropped_ = dropout(contextDropped_, keywords::mask = dropMaskContext_);

    if(layerNorm_) {
      gammaContext_ = graph->param(prefix + "_att_gamma1", {1, dimEncState}, keywords::init = inits::from_value(1.0));
      gammaState_ = graph->param(prefix + "_att_gamma2", {1, dimEncState}, keywords::init = inits::from_value(1.0));
      mappedContext_ = layer_norm(dot(contextDropped_, Ua_), gammaContext_, ba_);
    } else {
      mappedContext_ = affine(contextDropped_, Ua_, ba_);
    }

    auto softmaxMask = encState_->getMask();
    if(softmaxMask) {
      Shape shape = {softmaxMask->shape()[-3], softmaxMask->shape()[-2]};
      softmaxMask_ = transpose(reshape(softmaxMask, shape));
    }
  }

  Expr apply(State state) {
    using namespace keywords;
    auto recState = state.output;

    int dimBatch = contextDropped_->shape()[-2];
    int srcWords = contextDropped_->shape()[-3];
    int dimBeam = 1;
    if(recState->shape().size() > 3)
      dimBeam = recState->shape()[-4];

    if(dropMaskState_)
      recState = dropout(recState, keywords::mask = dropMaskState_);

    auto mappedState = dot(recState, Wa_);
    if(layerNorm_)
      mappedState = layer_norm(mappedState, gammaState_);

    auto attReduce = attOps(va_, mappedContext_, mappedState);

    // softmax over source positions, masked by the source mask
    auto e = reshape(transpose(softmax(transpose(attReduce), softmaxMask_)), {dimBeam, srcWords, dimBatch, 1});
    auto alignedSource = scalar_product(encState_->getAttended(), e, axis = -3);

    contexts_.push_back(alignedSource);
    alignments_.push_back(e);
    return alignedSource;
  }

  std::vector<Expr>& getContexts() { return contexts_; }
  Expr getContext() { return concatenate(contexts_, keywords::axis = -3); }
  std::vector<Expr>& getAlignments() { return alignments_; }

  virtual void clear() {
    contexts_.clear();
    alignments_.clear();
  }
  int dimOutput() { return encState_->getContext()->shape()[-1]; }
};
using Attention = GlobalAttention;

// ---------------------------------------------------------------------------
// time loop and layer stack.   reference: rnn.h:54-255
// ---------------------------------------------------------------------------
class SingleLayerRNN {
private:
  Ptr<Cell> cell_;
  dir direction_;
  States last_;

  States apply(const Expr input, const States initialState, const Expr mask = nullptr) {
    last_.clear();
    State state = initialState.front();
    cell_->clear();

    // the input projection of all time steps is ONE large GEMM
    auto xWs = cell_->applyInput({input});

    size_t timeSteps = input->shape()[-3];
    States outputs;
    for(size_t i = 0; i < timeSteps; ++i) {
      int j = (int)i;
      if(direction_ == dir::backward)
        j = (int)(timeSteps - i - 1);

      std::vector<Expr> steps(xWs.size());
      std::transform(xWs.begin(), xWs.end(), steps.begin(), [j](Expr e) { return step(e, j, -3); });

      if(mask)
        state = cell_->applyState(steps, state, step(mask, j, -3));
      else
        state = cell_->applyState(steps, state);
      outputs.push_back(state);
    }
    if(direction_ == dir::backward)
      outputs.reverse();
    last_.push_back(outputs.back());
    return outputs;
  }

  States apply(const Expr input, const Expr mask = nullptr) {
    auto graph = input->graph();
    int dimBatch = input->shape()[-2];
    int dimState = cell_->getOptions()->get<int>("dimState");
    auto output = graph->zeros({1, dimBatch, dimState});
    State startState{output, output};
    return apply(input, States({startState}), mask);
  }

public:
  SingleLayerRNN(Ptr<Options> options) : direction_((dir)options->get<int>("direction", (int)dir::forward)) {}

  Expr transduce(Expr input, Expr mask = nullptr) { return apply(input, mask).outputs(); }
  Expr transduce(Expr input, States states, Expr mask = nullptr) { return apply(input, states, mask).outputs(); }
  Expr transduce(Expr input, State state, Expr mask = nullptr) { return apply(input, States({state}), mask).outputs(); }

  States lastCellStates() { return last_; }
  void push_back(Ptr<Cell> cell) { cell_ = cell; }
  Ptr<Cell> at(int i) {
    ABORT_IF(i > 0, "SingleRNN only has one cell");
    return cell_;
  }
};

class RNN : public std::enable_shared_from_this<RNN> {
private:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
  bool skip_;
  bool skipFirst_;
  std::vector<Ptr<SingleLayerRNN>> rnns_;

  template <class Transduce>
  Expr run(Expr input, Transduce transduceLayer) {
    ABORT_IF(rnns_.empty(), "0 layers in RNN");
    Expr output;
    Expr layerInput = input;
    for(size_t i = 0; i < rnns_.size(); ++i) {
      Expr lazyInput = layerInput;
      auto cell = rnns_[i]->at(0);
      auto lazyInputs = cell->getLazyInputs(shared_from_this());
      if(!lazyInputs.empty()) {
        lazyInputs.push_back(layerInput);
        lazyInput = concatenate(lazyInputs, keywords::axis = -1);
      }
      auto layerOutput = transduceLayer(i, lazyInput);
      if(skip_ && (skipFirst_ || i > 0))
        output = layerOutput + layerInput;
      else
        output = layerOutput;
      layerInput = output;
    }
    return output;
  }

public:
  RNN(Ptr<ExpressionGraph> graph, Ptr<Options> options)
      : graph_(graph), options_(options), skip_(options->get<bool>("skip", false)), skipFirst_(options->get<bool>("skipFirst", false)) {}

  void push_back(Ptr<Cell> cell) {
    auto rnn = New<SingleLayerRNN>(cell->getOptions());
    rnn->push_back(cell);
    rnns_.push_back(rnn);
  }

  Expr transduce(Expr input, Expr mask = nullptr) {
    return run(input, [&](size_t i, Expr in) { return rnns_[i]->transduce(in, mask); });
  }
  Expr transduce(Expr input, States states, Expr mask = nullptr) {
    return run(input, [&](size_t i, Expr in) { return rnns_[i]->transduce(in, States({states[i]}), mask); });
  }
  Expr transduce(Expr input, State state, Expr mask = nullptr) {
    return run(input, [&](size_t i, Expr in) { return rnns_[i]->transduce(in, States({state}), mask); });
  }

  States lastCellStates() {
    States temp;
    for(auto rnn : rnns_)
      temp.push_back(rnn->lastCellStates().back());
    return temp;
  }
  Ptr<Cell> at(int i) { return rnns_[i]->at(0); }
  Ptr<Options> getOptions() { return options_; }
};

// ---------------------------------------------------------------------------
// builders:  rnn::rnn(graph)("type","gru")("dimState",1024).push_back(rnn::stacked_cell(graph)...)
// (the reference's Accumulator<Factory> chain, constructors.h:10-221)
// ---------------------------------------------------------------------------
class StackableFactory {
protected:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;

public:
  StackableFactory(Ptr<ExpressionGraph> graph) : graph_(graph), options_(New<Options>()) {}
  virtual ~StackableFactory() {}
  Ptr<Options> getOptions() { return options_; }
  virtual bool isCell() const = 0;
};

template <class Derived, class Base>
class Fluent : public Base {
public:
  using Base::Base;
  template <typename T>
  Derived& operator()(const std::string& key, T value) {
    this->options_->set(key, value);
    return static_cast<Derived&>(*this);
  }
  Derived& operator()(const std::string& key, dir value) {
    this->options_->set(key, (int)value);
    return static_cast<Derived&>(*this);
  }
};

class CellFactory : public StackableFactory {
protected:
  std::vector<std::function<Expr(Ptr<rnn::RNN>)>> inputs_;

public:
  using StackableFactory::StackableFactory;
  bool isCell() const { return true; }

  virtual Ptr<Cell> construct() {
    std::string type = options_->get<std::string>("type");
    Ptr<Cell> cell;
    if(type == "gru")
      cell = New<GRU>(graph_, options_);
    else if(type == "lstm")
      cell = New<LSTM>(graph_, options_);
    else if(type == "tanh")
      cell = New<Tanh>(graph_, options_);
    else
      ABORT("Unknown RNN cell type:", type);
    cell->setLazyInputs(inputs_);
    return cell;
  }
  virtual Ptr<CellFactory> cloneFactory() const { return Ptr<CellFactory>(new CellFactory(copyOf(*this))); }

  void add_input(std::function<Expr(Ptr<rnn::RNN>)> func) { inputs_.push_back(func); }
  void add_input(Expr input) {
    inputs_.push_back([input](Ptr<rnn::RNN>) { return input; });
  }

protected:
  // deep copy of the option bag (factories are value-copied into their parents)
  template <class F>
  static F copyOf(const F& f) {
    F c(f);
    c.options_ = f.options_->clone();
    return c;
  }
};
class cell : public Fluent<cell, CellFactory> {
public:
  using Fluent::Fluent;
  Ptr<CellFactory> cloneFactory() const { return Ptr<CellFactory>(new cell(copyOf(*this))); }
};

class InputFactory : public StackableFactory {
public:
  using StackableFactory::StackableFactory;
  bool isCell() const { return false; }
  virtual Ptr<CellInput> construct() = 0;
  virtual Ptr<InputFactory> cloneFactory() const = 0;
};

class AttentionFactory : public InputFactory {
protected:
  Ptr<EncoderState> state_;

public:
  using InputFactory::InputFactory;
  Ptr<CellInput> construct() {
    ABORT_IF(!state_, "EncoderState not set");
    return New<Attention>(graph_, options_, state_);
  }
};
class attention : public Fluent<attention, AttentionFactory> {
public:
  using Fluent::Fluent;
  attention& set_state(Ptr<EncoderState> state) {
    state_ = state;
    return *this;
  }
  Ptr<InputFactory> cloneFactory() const {
    attention c(*this);
    c.options_ = options_->clone();
    return Ptr<InputFactory>(new attention(c));
  }
};

class StackedCellFactory : public CellFactory {
protected:
  std::vector<Ptr<StackableFactory>> stackableFactories_;

public:
  using CellFactory::CellFactory;

  // reference: constructors.h:96-126
  Ptr<Cell> construct() {
    auto stacked = New<StackedCell>(graph_, options_);
    int lastDimInput = options_->get<int>("dimInput");
    for(size_t i = 0; i < stackableFactories_.size(); ++i) {
      auto sf = stackableFactories_[i];
      if(sf->isCell()) {
        auto cellFactory = std::dynamic_pointer_cast<CellFactory>(sf);
        cellFactory->getOptions()->merge(options_);
        sf->getOptions()->set("dimInput", lastDimInput);
        lastDimInput = 0;
        if(i == 0)
          for(auto f : inputs_)
            cellFactory->add_input(f);
        stacked->push_back(cellFactory->construct());
      } else {
        auto inputFactory = std::dynamic_pointer_cast<InputFactory>(sf);
        inputFactory->getOptions()->merge(options_);
        auto input = inputFactory->construct();
        stacked->push_back(input);
        lastDimInput += input->dimOutput();
      }
    }
    return stacked;
  }
};
class stacked_cell : public Fluent<stacked_cell, StackedCellFactory> {
public:
  using Fluent::Fluent;
  stacked_cell& push_back(const CellFactory& f) {
    stackableFactories_.push_back(f.cloneFactory());
    return *this;
  }
  stacked_cell& push_back(const attention& f) {
    stackableFactories_.push_back(f.cloneFactory());
    return *this;
  }
  Ptr<CellFactory> cloneFactory() const {
    stacked_cell c(*this);
    c.options_ = options_->clone();
    return Ptr<CellFactory>(new stacked_cell(c));
  }
};

class rnn {
protected:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
  std::vector<Ptr<CellFactory>> layerFactories_;

public:
  explicit rnn(Ptr<ExpressionGraph> graph) : graph_(graph), options_(New<Options>()) {}

  template <typename T>
  rnn& operator()(const std::string& key, T value) {
    options_->set(key, value);
    return *this;
  }
  rnn& operator()(const std::string& key, dir value) {
    options_->set(key, (int)value);
    return *this;
  }

  rnn& push_back(const CellFactory& f) {
    layerFactories_.push_back(f.cloneFactory());
    return *this;
  }

  // reference: constructors.h:157-198
  Ptr<RNN> construct() {
    auto r = New<RNN>(graph_, options_);
    dir direction = (dir)options_->get<int>("direction", (int)dir::forward);
    for(size_t i = 0; i < layerFactories_.size(); ++i) {
      auto lf = layerFactories_[i];
      lf->getOptions()->merge(options_);
      if(i > 0) {
        int dimInput = layerFactories_[i - 1]->getOptions()->get<int>("dimState")
                       + lf->getOptions()->get<int>("dimInputExtra", 0);
        lf->getOptions()->set("dimInput", dimInput);
      }
      if(direction == dir::alternating_forward)
        lf->getOptions()->set("direction", (int)(i % 2 == 0 ? dir::forward : dir::backward));
      if(direction == dir::alternating_backward)
        lf->getOptions()->set("direction", (int)(i % 2 == 1 ? dir::forward : dir::backward));
      r->push_back(lf->construct());
    }
    return r;
  }
  Ptr<RNN> operator->() { return construct(); }
};

}  // namespace rnn
}  // namespace marian
