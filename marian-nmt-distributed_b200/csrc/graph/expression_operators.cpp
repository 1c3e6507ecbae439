// Node-building free functions (reference: src/graph/expression_operators.cu).
#include "graph/expression_operators.h"
#include "graph/node_operators_binary.h"
#include "graph/node_operators_unary.h"

namespace marian {

Expr debug(Expr a, const std::string& message) {
  a->debug(message);
  return a;
}

Expr logit(Expr a) { return Expression<LogitNodeOp>(a); }
Expr relu(Expr a) { return Expression<ReLUNodeOp>(a); }
Expr leakyrelu(Expr a) { return Expression<PReLUNodeOp>(0.01f, a); }
Expr prelu(Expr a, float alpha) { return Expression<PReLUNodeOp>(alpha, a); }
Expr swish(Expr a) { return Expression<SwishNodeOp>(a); }
Expr log(Expr a) { return Expression<LogNodeOp>(a); }
Expr exp(Expr a) { return Expression<ExpNodeOp>(a); }
Expr sqrt(Expr a, float eps) { return Expression<SqrtNodeOp>(a, eps); }
Expr square(Expr a) { return Expression<SquareNodeOp>(a); }
Expr tanh(const std::vector<Expr>& nodes) { return Expression<TanhNodeOp>(nodes); }

Expr operator-(Expr a) { return Expression<NegNodeOp>(a); }

Expr operator+(Expr a, Expr b) { return Expression<PlusNodeOp>(a, b); }
Expr operator-(Expr a, Expr b) { return Expression<MinusNodeOp>(a, b); }
Expr operator*(Expr a, Expr b) { return Expression<MultNodeOp>(a, b); }
Expr operator/(Expr a, Expr b) { return Expression<DivNodeOp>(a, b); }

Expr operator+(Expr a, float b) { return Expression<ScalarAddNodeOp>(a, b); }
Expr operator+(float a, Expr b) { return Expression<ScalarAddNodeOp>(b, a); }
Expr operator-(Expr a, float b) { return Expression<ScalarAddNodeOp>(a, -b); }
Expr operator-(float a, Expr b) { return Expression<ScalarAddNodeOp>(-b, a); }
Expr operator*(float a, Expr b) { return Expression<ScalarMultNodeOp>(b, a); }
Expr operator*(Expr a, float b) { return Expression<ScalarMultNodeOp>(a, b); }
Expr operator/(Expr a, float b) { return Expression<ScalarMultNodeOp>(a, 1.f / b); }

Expr dot(Expr a, Expr b, bool transA, bool transB, float scalar) {
  return Expression<DotNodeOp>(a, b, transA, transB, scalar);
}
Expr bdot(Expr a, Expr b, bool transA, bool transB, float scalar) {
  return Expression<DotBatchedNodeOp>(a, b, transA, transB, scalar);
}
Expr affine(Expr a, Expr b, Expr c) {
  std::vector<Expr> nodes = {a, b, c};
  return Expression<AffineNodeOp>(nodes);
}

Expr transpose(Expr a) {
  std::vector<int> axes(a->shape().size());
  for(size_t i = 0; i < axes.size(); ++i)
    axes[i] = (int)i;
  if(axes.size() > 1) {
    axes[axes.size() - 1] = (int)axes.size() - 2;
    axes[axes.size() - 2] = (int)axes.size() - 1;
  }
  return Expression<TransposeNodeOp>(a, axes);
}
Expr transpose(Expr a, const std::vector<int>& axes) { return Expression<TransposeNodeOp>(a, axes); }

Expr concatenate(const std::vector<Expr>& concats, keywords::axis_k ax) {
  return Expression<ConcatenateNodeOp>(concats, ax);
}
Expr repeat(Expr a, size_t repeats, keywords::axis_k ax) {
  if(repeats == 1)
    return a;
  return concatenate(std::vector<Expr>(repeats, a), ax);
}

Expr reshape(Expr a, Shape shape) { return Expression<ReshapeNodeOp>(a, shape); }
Expr atleast_1d(Expr a) { return atleast_nd(a, 1); }
Expr atleast_2d(Expr a) { return atleast_nd(a, 2); }
Expr atleast_3d(Expr a) { return atleast_nd(a, 3); }
Expr atleast_4d(Expr a) { return atleast_nd(a, 4); }
Expr atleast_nd(Expr a, size_t dims) {
  if(a->shape().size() >= dims)
    return a;
  Shape nShape;
  nShape.resize(dims);
  for(int i = 1; i <= (int)a->shape().size(); ++i)
    nShape.set(-i, a->shape()[-i]);
  return reshape(a, nShape);
}
Expr flatten(Expr a) {
  Shape shape = {a->shape().elements()};
  return Expression<ReshapeNodeOp>(a, shape);
}
Expr flatten_2d(Expr a) {
  Shape shape = {a->shape().elements() / a->shape()[-1], a->shape()[-1]};
  return Expression<ReshapeNodeOp>(a, shape);
}

Expr rows(Expr a, const std::vector<size_t>& indices) { return Expression<RowsNodeOp>(a, indices); }
Expr rows(Expr a, const std::vector<size_t>& indices, ExpressionGraph::BatchFillI fill, Ptr<data::CorpusBatch> batch) {
  return Expression<RowsNodeOp>(a, indices, fill, batch);
}
Expr cols(Expr a, const std::vector<size_t>& indices) { return Expression<ColsNodeOp>(a, indices); }
Expr select(Expr a, int axis, const std::vector<size_t>& indices) {
  return Expression<SelectNodeOp>(a, axis, indices);
}

Expr sum(Expr a, keywords::axis_k ax) { return Expression<SumNodeOp>(a, ax); }
Expr mean(Expr a, keywords::axis_k ax) { return Expression<MeanNodeOp>(a, ax); }
Expr softmax(Expr a, Expr mask) { return Expression<SoftmaxNodeOp>(a, mask); }
Expr logsoftmax(Expr a) { return Expression<LogSoftmaxNodeOp>(a); }
Expr cross_entropy(Expr a, Expr b) { return Expression<CrossEntropyNodeOp>(a, b); }
Expr scalar_product(Expr a, Expr b, keywords::axis_k ax) { return Expression<ScalarProductNodeOp>(a, b, ax); }
Expr weighted_average(Expr in, Expr weights, keywords::axis_k ax) {
  auto p = scalar_product(in, weights, ax);
  auto s = sum(weights, ax);
  return p / s;
}

Expr step(Expr a, int step, int axis) { return Expression<StepNodeOp>(a, step, axis); }
Expr shift(Expr a, Shape shift) { return Expression<ShiftNodeOp>(a, shift); }

Expr layer_norm(Expr x, Expr gamma, Expr beta, float eps) {
  std::vector<Expr> nodes = {x, gamma};
  if(beta)
    nodes.push_back(beta);
  return Expression<LayerNormalizationOp>(nodes, eps);
}
Expr residual_layer_norm(Expr x, Expr residual, Expr gamma, Expr beta, float eps) {
  std::vector<Expr> nodes = {x, residual, gamma, beta};
  return Expression<ResidualLayerNormOp>(nodes, eps);
}
Expr highway(Expr y, Expr x, Expr t) {
  std::vector<Expr> nodes = {y, x, t};
  return Expression<HighwayNodeOp>(nodes);
}
Expr multi_head_attention(Expr q, Expr k, Expr v, Expr mask, int heads, float scale) {
  std::vector<Expr> nodes = {q, k, v};
  if(mask)
    nodes.push_back(mask);
  return Expression<MultiHeadAttentionNodeOp>(nodes, heads, scale);
}

}  // namespace marian
