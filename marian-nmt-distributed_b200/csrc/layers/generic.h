// Dense layers, MLP stacks, embeddings and the training cost.
//
// Parameter names, creation order and formulas follow the reference
// (src/layers/generic.h:44-225, src/layers/constructors.h:31-118,
// src/layers/generic.cpp:5-42), because creation order fixes the random
// initialisation stream (layers/param_initializers.h).  The YAML-backed
// Accumulator<Factory> machinery is replaced by one small fluent builder with
// the same call syntax:  mlp::dense(graph)("prefix", p)("dim", n)("activation", mlp::act::tanh)
#pragma once

#include <map>

#include "common/options.h"
#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "layers/param_initializers.h"

namespace marian {
namespace mlp {

enum struct act : int { linear, tanh, logit, ReLU, LeakyReLU, PReLU, swish };

class Dense {
public:
  Dense(Ptr<ExpressionGraph> graph, Ptr<Options> options) : graph_(graph), options_(options) {}

  void tie(const std::string& param, const std::string& tied) { tiedParams_[param] = graph_->get(tied); }
  void tie_transposed(const std::string& param, const std::string& tied) {
    tiedParams_[param] = transpose(graph_->get(tied));
  }

  // several inputs: one affine per input, activation of the sum (reference: generic.h:63-135)
  Expr apply(const std::vector<Expr>& inputs) {
    ABORT_IF(inputs.empty(), "No inputs");
    if(inputs.size() == 1)
      return apply(inputs[0]);

    auto name = options_->get<std::string>("prefix");
    auto dim = options_->get<int>("dim");
    auto layerNorm = options_->get<bool>("layer-normalization", false);
    auto nematusNorm = options_->get<bool>("nematus-normalization", false);
    auto activation = (act)options_->get<int>("activation", (int)act::linear);

    std::vector<Expr> outputs;
    size_t i = 0;
    for(auto&& in : inputs) {
      std::string num = std::to_string(i);
      Expr W = tiedParams_.count("W" + num)
                   ? tiedParams_["W" + num]
                   : graph_->param(name + "_W" + num, {in->shape()[-1], dim}, keywords::init = inits::glorot_uniform);
      Expr b = tiedParams_.count("b" + num)
                   ? tiedParams_["b" + num]
                   : graph_->param(name + "_b" + num, {1, dim}, keywords::init = inits::zeros);
      if(layerNorm) {
        if(nematusNorm) {
          auto ln_s = graph_->param(name + "_ln_s" + num, {1, dim}, keywords::init = inits::from_value(1.f));
          auto ln_b = graph_->param(name + "_ln_b" + num, {1, dim}, keywords::init = inits::zeros);
          outputs.push_back(layer_norm(affine(in, W, b), ln_s, ln_b, NEMATUS_LN_EPS));
        } else {
          auto gamma = graph_->param(name + "_gamma" + num, {1, dim}, keywords::init = inits::from_value(1.0));
          outputs.push_back(layer_norm(dot(in, W), gamma, b));
        }
      } else {
        outputs.push_back(affine(in, W, b));
      }
      i++;
    }
    // the reference only implements the n-ary form for tanh (expression_operators.cu:258-290)
    ABORT_IF(activation != act::tanh, "Multi-input dense layers support only the tanh activation");
    return tanh(outputs);
  }

  // reference: generic.h:137-196
  Expr apply(Expr input) {
    auto name = options_->get<std::string>("prefix");
    auto dim = options_->get<int>("dim");
    auto layerNorm = options_->get<bool>("layer-normalization", false);
    auto nematusNorm = options_->get<bool>("nematus-normalization", false);
    auto activation = (act)options_->get<int>("activation", (int)act::linear);

    Expr W = tiedParams_.count("W")
                 ? tiedParams_["W"]
                 : graph_->param(name + "_W", {input->shape()[-1], dim}, keywords::init = inits::glorot_uniform);
    Expr b = tiedParams_.count("b") ? tiedParams_["b"]
                                    : graph_->param(name + "_b", {1, dim}, keywords::init = inits::zeros);

    Expr out;
    if(layerNorm) {
      if(nematusNorm) {
        auto ln_s = graph_->param(name + "_ln_s", {1, dim}, keywords::init = inits::from_value(1.f));
        auto ln_b = graph_->param(name + "_ln_b", {1, dim}, keywords::init = inits::zeros);
        out = layer_norm(affine(input, W, b), ln_s, ln_b, NEMATUS_LN_EPS);
      } else {
        auto gamma = graph_->param(name + "_gamma", {1, dim}, keywords::init = inits::from_value(1.0));
        out = layer_norm(dot(input, W), gamma, b);
      }
    } else {
      out = affine(input, W, b);
    }

    switch(activation) {
      case act::linear: return out;
      case act::tanh: return tanh(out);
      case act::logit: return logit(out);
      case act::ReLU: return relu(out);
      case act::LeakyReLU: return leakyrelu(out);
      case act::PReLU: return prelu(out);
      case act::swish: return swish(out);
      default: return out;
    }
  }

private:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
  std::map<std::string, Expr> tiedParams_;
};

// fluent builder: dense(graph)("prefix", ..)("dim", ..).tie_transposed("W", name)
class dense {
public:
  explicit dense(Ptr<ExpressionGraph> graph) : graph_(graph), options_(New<Options>()) {}

  template <typename T>
  dense& operator()(const std::string& key, T value) {
    options_->set(key, value);
    return *this;
  }
  dense& operator()(const std::string& key, act value) {
    options_->set(key, (int)value);
    return *this;
  }
  dense& tie(const std::string& param, const std::string& tied) {
    tied_.push_back({param, tied});
    return *this;
  }
  dense& tie_transposed(const std::string& param, const std::string& tied) {
    tiedTransposed_.push_back({param, tied});
    return *this;
  }
  Ptr<Options> getOptions() { return options_; }

  Ptr<Dense> construct() {
    auto d = New<Dense>(graph_, options_);
    for(auto& p : tied_)
      d->tie(p.first, p.second);
    for(auto& p : tiedTransposed_)
      d->tie_transposed(p.first, p.second);
    return d;
  }

private:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
  std::vector<std::pair<std::string, std::string>> tied_, tiedTransposed_;
};

class MLP {
public:
  template <typename... Args>
  Expr apply(Args... args) {
    std::vector<Expr> av = {args...};
    Expr output = av.size() == 1 ? layers_[0]->apply(av[0]) : layers_[0]->apply(av);
    for(size_t i = 1; i < layers_.size(); ++i)
      output = layers_[i]->apply(output);
    return output;
  }
  void push_back(Ptr<Dense> layer) { layers_.push_back(layer); }

private:
  std::vector<Ptr<Dense>> layers_;
};

// mlp(graph).push_back(dense(...)).push_back(dense(...))->apply(x, y)
class mlp {
public:
  explicit mlp(Ptr<ExpressionGraph> graph) : graph_(graph), options_(New<Options>()) {}
  template <typename T>
  mlp& operator()(const std::string& key, T value) {
    options_->set(key, value);
    return *this;
  }
  mlp& push_back(const dense& d) {
    layers_.push_back(d);
    return *this;
  }
  Ptr<MLP> construct() {
    auto m = New<MLP>();
    for(auto& layer : layers_) {
      layer.getOptions()->merge(*options_);
      m->push_back(layer.construct());
    }
    return m;
  }
  Ptr<MLP> operator->() { return construct(); }

private:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
  std::vector<dense> layers_;
};

}  // namespace mlp

// embedding(graph)("dimVocab", V)("dimEmb", d)("prefix", name).construct()
// reference: generic.h:201-225
class embedding {
public:
  explicit embedding(Ptr<ExpressionGraph> graph) : graph_(graph), options_(New<Options>()) {}
  template <typename T>
  embedding& operator()(const std::string& key, T value) {
    options_->set(key, value);
    return *this;
  }
  Expr construct() {
    std::string name = options_->get<std::string>("prefix");
    int dimVoc = options_->get<int>("dimVocab");
    int dimEmb = options_->get<int>("dimEmb");
    bool fixed = options_->get<bool>("fixed", false);
    return graph_->param(name, {dimVoc, dimEmb}, keywords::init = inits::glorot_uniform, keywords::fixed = fixed);
  }

private:
  Ptr<ExpressionGraph> graph_;
  Ptr<Options> options_;
};

Expr Cost(Expr logits, Expr indices, Expr mask, std::string costType = "cross-entropy", float smoothing = 0);

}  // namespace marian
