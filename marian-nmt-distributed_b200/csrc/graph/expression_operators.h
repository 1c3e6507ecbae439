// Free functions that build graph nodes: the operator API models are written in.
// Same names and argument meaning as the reference's
// src/graph/expression_operators.h:6-157.
#pragma once

#include "graph/expression_graph.h"

namespace marian {

Expr debug(Expr a, const std::string& message = "");

Expr logit(Expr a);
Expr relu(Expr a);
Expr leakyrelu(Expr a);
Expr prelu(Expr a, float alpha = 0.01f);
Expr swish(Expr a);
Expr log(Expr a);
Expr exp(Expr a);
Expr sqrt(Expr a, float eps = 0.f);
Expr square(Expr a);
Expr tanh(const std::vector<Expr>&);
template <typename... Args>
Expr tanh(Args... args) {
  std::vector<Expr> nodes{args...};
  return tanh(nodes);
}

Expr operator-(Expr a);

Expr operator+(Expr a, Expr b);
Expr operator-(Expr a, Expr b);
Expr operator*(Expr a, Expr b);
Expr operator/(Expr a, Expr b);

Expr operator+(float a, Expr b);
Expr operator+(Expr a, float b);
Expr operator-(float a, Expr b);
Expr operator-(Expr a, float b);
Expr operator*(float a, Expr b);
Expr operator*(Expr a, float b);
Expr operator/(Expr a, float b);

Expr dot(Expr a, Expr b, bool transA = false, bool transB = false, float scalar = 1.f);
Expr bdot(Expr a, Expr b, bool transA = false, bool transB = false, float scalar = 1.f);
Expr affine(Expr a, Expr b, Expr c);

Expr transpose(Expr a);
Expr transpose(Expr a, const std::vector<int>& axes);

Expr concatenate(const std::vector<Expr>& concats, keywords::axis_k ax = keywords::axis_k{0});
Expr repeat(Expr a, size_t repeats, keywords::axis_k ax = keywords::axis_k{0});

Expr reshape(Expr a, Shape shape);
Expr atleast_1d(Expr a);
Expr atleast_2d(Expr a);
Expr atleast_3d(Expr a);
Expr atleast_4d(Expr a);
Expr atleast_nd(Expr a, size_t dims);
Expr flatten(Expr a);
Expr flatten_2d(Expr a);

Expr rows(Expr a, const std::vector<size_t>& indices);
// rows() whose indices are a function of the current batch (replayable upload)
Expr rows(Expr a, const std::vector<size_t>& indices, ExpressionGraph::BatchFillI fill, Ptr<data::CorpusBatch> batch);
Expr cols(Expr a, const std::vector<size_t>& indices);
Expr select(Expr a, int axis, const std::vector<size_t>& indices);

Expr sum(Expr a, keywords::axis_k ax = keywords::axis_k{0});
Expr mean(Expr a, keywords::axis_k ax = keywords::axis_k{0});
Expr softmax(Expr a, Expr mask = nullptr);
Expr logsoftmax(Expr a);
Expr cross_entropy(Expr a, Expr b);
Expr scalar_product(Expr a, Expr b, keywords::axis_k ax = keywords::axis_k{0});
Expr weighted_average(Expr in, Expr weights, keywords::axis_k ax = keywords::axis_k{0});

Expr step(Expr a, int step, int axis);
Expr shift(Expr a, Shape shift);

Expr layer_norm(Expr x, Expr gamma, Expr beta = nullptr, float eps = 1e-9);
// layer_norm(x + residual, gamma, beta) as one operator (see ResidualLayerNormOp)
Expr residual_layer_norm(Expr x, Expr residual, Expr gamma, Expr beta, float eps = 1e-9);
Expr highway(Expr y, Expr x, Expr t);
// fused multi-head attention core on [beam, B, T, d] projections; mask additive (may be null)
Expr multi_head_attention(Expr q, Expr k, Expr v, Expr mask, int heads, float scale);

// inverted dropout with an explicit mask node (reference: expression_operators.h:122-133)
template <typename... Args>
Expr dropout(Expr x, Args... args) {
  auto mask = keywords::Get(keywords::mask, Expr(nullptr), args...);
  float dropout_prob = keywords::Get(keywords::dropout_prob, 0.0f, args...);
  ABORT_IF(!mask && !dropout_prob, "Neither mask nor dropout prob given");
  if(!mask)
    mask = x->graph()->dropout(dropout_prob, x->shape());
  return x * mask;
}

}  // namespace marian
