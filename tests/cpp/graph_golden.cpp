// Golden-vector test driver: rebuilds, with THIS repo's graph API, the graphs of
// the reference's own unit tests and returns the values those tests assert on.
//   /root/reference/src/tests/operator_tests.cpp:19-293
//   /root/reference/src/tests/rnn_tests.cpp:32-250
//   /root/reference/src/tests/attention_tests.cpp:32-105
// The expected numbers live in tests/golden/reference_unit_tests.json (extracted
// from those files by tests/golden/extract_reference_goldens.py).  The same
// source is linked against the CPU oracle (pins the oracle) and against the
// CUDA library (`-m gpu` tests), so it reads like the reference's tests while
// exercising whichever backend is loaded.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

#include "marian_b200.h"

#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "models/states.h"
#include "rnn/rnn.h"

using namespace marian;

namespace {

typedef std::vector<float> Values;

Ptr<ExpressionGraph> newGraph() {
  auto graph = New<ExpressionGraph>();
  graph->setDevice(0);
  graph->reserveWorkspaceMB(16);
  return graph;
}

void append(Values& out, Expr e) {
  Values v;
  e->val()->get(v);
  out.insert(out.end(), v.begin(), v.end());
}

const std::vector<size_t> vWords = {43, 2, 83, 78, 6, 38, 80, 40, 40, 70, 26, 60, 106, 13, 111, 32,
                                    126, 62, 115, 72, 127, 82, 55, 0, 86, 0, 124, 0, 0, 0, 0, 0};
const std::vector<size_t> vMask = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                   1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 1, 0};

// ---- operator_tests.cpp -------------------------------------------------
Values opDot() {
  auto graph = newGraph();
  Values vA({1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}), vB({1, 2, 3, 4, 5, 6});
  auto A = graph->param("A", {2, 2, 3}, keywords::init = inits::from_vector(vA));
  auto B = graph->param("B", {3, 2}, keywords::init = inits::from_vector(vB));
  auto C = dot(A, B);
  graph->forward();
  ABORT_IF(C->shape() != Shape({2, 2, 2}), "dot: wrong shape");
  Values out;
  append(out, C);
  return out;
}

Values opScalarMult() {
  auto graph = newGraph();
  Values vB({1, 2, 3, 4, 5, 6});
  auto B = graph->param("B", {3, 2}, keywords::init = inits::from_vector(vB));
  auto B2 = B * 2.0f;
  graph->forward();
  Values out;
  append(out, B2);
  return out;
}

Values opSoftmax() {
  auto graph = newGraph();
  Values in({-.2f, -.3f, 4.5f, 5.2f, -10.f, 101.45f, -100.05f, 1.05e-5f});
  auto input = graph->constant({2, 2, 2}, keywords::init = inits::from_vector(in));
  auto sm = softmax(input);
  auto lsm = logsoftmax(input);
  graph->forward();
  Values out;
  append(out, sm);
  append(out, lsm);
  return out;
}

Values opBroadcast() {
  auto graph = newGraph();
  Values vA({1, -2, 3, -4}), vB({0.5, 1.5});
  auto a = graph->constant({2, 2, 1}, keywords::init = inits::from_vector(vA));
  auto b = graph->constant({2, 1}, keywords::init = inits::from_vector(vB));
  auto add = a + b;
  auto minus = b - a;
  auto mult = a * b;
  auto div = a / b;
  graph->forward();
  ABORT_IF(add->shape() != Shape({2, 2, 1}), "broadcast: wrong shape");
  Values out;
  append(out, add);
  append(out, minus);
  append(out, mult);
  append(out, div);
  return out;
}

Values opTranspose() {
  auto graph = newGraph();
  Values vA({1, 2, 3, 4, 5, 6, 7, 8});
  auto a = graph->constant({2, 4}, keywords::init = inits::from_vector(vA));
  auto t1 = transpose(a);
  auto t2 = transpose(t1);
  auto t3 = transpose(reshape(t1, {2, 2, 2}));
  auto t4 = transpose(reshape(a, {2, 1, 2, 2}), {1, 3, 2, 0});
  auto t5 = transpose(reshape(a, {2, 1, 2, 2}), {2, 0, 1, 3});
  graph->forward();
  ABORT_IF(t1->shape() != Shape({4, 2}), "t1 shape");
  ABORT_IF(t4->shape() != Shape({1, 2, 2, 2}), "t4 shape");
  ABORT_IF(t5->shape() != Shape({2, 2, 1, 2}), "t5 shape");
  Values out;
  append(out, t1);
  append(out, t2);
  append(out, t3);
  append(out, t4);
  append(out, t5);
  return out;
}

Values opReductions() {
  auto graph = newGraph();
  Values vA({1, 2, 3, 4, 5, 6, 7, 8});
  auto a = graph->constant({2, 4}, keywords::init = inits::from_vector(vA));
  auto s1 = sum(a, keywords::axis = 0);
  auto s2 = sum(a, keywords::axis = 1);
  auto m3 = mean(s1, keywords::axis = 1);
  auto sp = scalar_product(s2, s2, keywords::axis = 0);
  auto wa = weighted_average(a, s1, keywords::axis = -1);
  graph->forward();
  ABORT_IF(s1->shape() != Shape({1, 4}), "s1 shape");
  ABORT_IF(s2->shape() != Shape({2, 1}), "s2 shape");
  ABORT_IF(wa->shape() != Shape({2, 1}), "wa shape");
  Values out;
  append(out, s1);
  append(out, s2);
  append(out, m3);
  append(out, sp);
  append(out, wa);
  return out;
}

Values opConcat() {
  auto graph = newGraph();
  auto in1 = graph->constant({1, 2, 2, 3}, keywords::init = inits::from_value(1));
  auto in2 = graph->constant({1, 2, 2, 3}, keywords::init = inits::from_value(2));
  auto in3 = graph->constant({1, 2, 2, 3}, keywords::init = inits::from_value(3));
  auto in4 = graph->constant({1, 2, 2, 3}, keywords::init = inits::from_value(4));
  auto c1 = concatenate({in1, in2, in3, in4}, keywords::axis = 2);
  auto c2 = concatenate({in1, in2, in3, in4}, keywords::axis = -1);
  auto c3 = concatenate({in1, in2, in3, in4}, keywords::axis = -3);
  auto c4 = concatenate({in1, in2, in3, in4}, keywords::axis = 0);
  graph->forward();
  ABORT_IF(c1->shape() != Shape({1, 2, 8, 3}), "c1 shape");
  ABORT_IF(c2->shape() != Shape({1, 2, 2, 12}), "c2 shape");
  ABORT_IF(c3->shape() != Shape({1, 8, 2, 3}), "c3 shape");
  ABORT_IF(c4->shape() != Shape({4, 2, 2, 3}), "c4 shape");
  Values out;
  append(out, c1);
  append(out, c2);
  append(out, c3);
  append(out, c4);
  return out;
}

Values opLayerNorm() {
  auto graph = newGraph();
  Config::seed = 1234;
  auto a = graph->constant({2, 2, 4}, keywords::init = inits::glorot_uniform);
  auto gamma = graph->param("gamma", {1, 4}, keywords::init = inits::ones);
  auto beta = graph->param("beta", {1, 4}, keywords::init = inits::zeros);
  auto ln = layer_norm(a, gamma, beta);
  graph->forward();
  Values out;
  append(out, ln);
  return out;
}

// ---- rnn_tests.cpp ------------------------------------------------------
Values rnnSimple() {
  Config::seed = 1234;
  auto graph = newGraph();
  auto input = graph->constant({4, 1, 4}, keywords::init = inits::glorot_uniform);
  auto rnn = rnn::rnn(graph)("prefix", "rnntest")("type", "tanh")("dimInput", 4)("dimState", 4)
                 .push_back(rnn::cell(graph))
                 .construct();
  auto output = rnn->transduce(input);
  graph->forward();
  ABORT_IF(output->shape() != Shape({4, 1, 4}), "rnn output shape");
  Values out;
  append(out, output);
  return out;
}

Expr buildRnn(Ptr<ExpressionGraph> graph,
              std::string prefix,
              Expr input,
              Expr mask,
              int dimRnn = 32,
              int depth = 1,
              int cellDepth = 1,
              std::string type = "bidirectional",
              std::string cellType = "gru",
              bool layerNorm = false,
              bool skip = false) {
  using namespace keywords;
  int dimEmb = input->shape()[-1];
  int first, second;
  if(type == "bidirectional" || type == "alternating") {
    first = depth;
    second = 0;
  } else {
    first = 1;
    second = depth - first;
  }
  auto forward = type == "alternating" ? rnn::dir::alternating_forward : rnn::dir::forward;
  auto backward = type == "alternating" ? rnn::dir::alternating_backward : rnn::dir::backward;

  auto rnnFw = rnn::rnn(graph)("type", cellType)("direction", forward)("dimInput", dimEmb)("dimState", dimRnn)(
      "layer-normalization", layerNorm)("skip", skip);
  for(int i = 1; i <= first; ++i) {
    auto stacked = rnn::stacked_cell(graph);
    for(int j = 1; j <= cellDepth; ++j) {
      std::string paramPrefix = prefix + "_bi";
      if(i > 1)
        paramPrefix += "_l" + std::to_string(i);
      if(i > 1 || j > 1)
        paramPrefix += "_cell" + std::to_string(j);
      stacked.push_back(rnn::cell(graph)("prefix", paramPrefix));
    }
    rnnFw.push_back(stacked);
  }

  auto rnnBw = rnn::rnn(graph)("type", cellType)("direction", backward)("dimInput", dimEmb)("dimState", dimRnn)(
      "layer-normalization", layerNorm)("skip", skip);
  for(int i = 1; i <= first; ++i) {
    auto stacked = rnn::stacked_cell(graph);
    for(int j = 1; j <= cellDepth; ++j) {
      std::string paramPrefix = prefix + "_bi_r";
      if(i > 1)
        paramPrefix += "_l" + std::to_string(i);
      if(i > 1 || j > 1)
        paramPrefix += "_cell" + std::to_string(j);
      stacked.push_back(rnn::cell(graph)("prefix", paramPrefix));
    }
    rnnBw.push_back(stacked);
  }

  auto fw = rnnFw->transduce(input, mask);
  auto bw = rnnBw->transduce(input, mask);
  auto context = concatenate({fw, bw}, axis = (int)input->shape().size() - 1);

  if(second > 0) {
    auto rnnUni = rnn::rnn(graph)("type", cellType)("dimInput", 2 * dimRnn)("dimState", dimRnn)(
        "layer-normalization", layerNorm)("skip", skip);
    for(int i = first + 1; i <= second + first; ++i) {
      auto stacked = rnn::stacked_cell(graph);
      for(int j = 1; j <= cellDepth; ++j) {
        std::string paramPrefix = prefix + "_l" + std::to_string(i) + "_cell" + std::to_string(j);
        stacked.push_back(rnn::cell(graph)("prefix", paramPrefix));
      }
      rnnUni.push_back(stacked);
    }
    context = rnnUni->transduce(context);
  }
  return context;
}

Values rnnS2SEncoder() {
  Config::seed = 1234;
  auto graph = newGraph();
  int dimEmb = 16, dimBatch = 4, dimTime = 8;
  auto emb = graph->param("Embeddings", {128, dimEmb}, keywords::init = inits::glorot_uniform);
  auto input = reshape(rows(emb, vWords), {dimTime, dimBatch, dimEmb});
  auto mask = graph->constant({dimTime, dimBatch, 1}, keywords::init = inits::from_vector(vMask));

  int dimRnn = 32;
  auto context1 = buildRnn(graph, "enc1", input, mask, dimRnn);
  auto contextSum1 = sum(context1, keywords::axis = 2);
  auto context2 = buildRnn(graph, "enc2", input, mask, dimRnn, 2, 2);
  auto contextSum2 = sum(context2, keywords::axis = 2);

  graph->forward();
  ABORT_IF(context1->shape() != Shape({dimTime, dimBatch, 2 * dimRnn}), "context1 shape");
  ABORT_IF(contextSum2->shape() != Shape({dimTime, dimBatch, 1}), "contextSum2 shape");
  Values out;
  append(out, contextSum1);
  append(out, contextSum2);
  return out;
}

// ---- attention_tests.cpp --------------------------------------------------
Values attentionContext() {
  Config::seed = 1234;
  auto graph = newGraph();
  int dimEmb = 16, dimBatch = 4, dimTime = 8;
  auto emb = graph->param("Embeddings", {128, dimEmb}, keywords::init = inits::glorot_uniform);
  auto input = reshape(rows(emb, vWords), {dimTime, dimBatch, dimEmb});
  auto mask = graph->constant({dimTime, dimBatch, 1}, keywords::init = inits::from_vector(vMask));

  auto rnn = rnn::rnn(graph)("prefix", "rnntest")("type", "gru")("dimInput", 16)("dimState", 8)
                 .push_back(rnn::cell(graph))
                 .construct();
  auto context = rnn->transduce(input, mask);
  auto encState = New<EncoderState>(context, mask, nullptr);

  auto options = New<Options>();
  options->set("dimState", 16);
  options->set("prefix", "rnntest_att");
  auto att = New<rnn::Attention>(graph, options, encState);

  std::vector<float> vState(64);
  int n = -32;
  std::generate(vState.begin(), vState.end(), [&n]() { return n++ / 64.f; });
  rnn::State state({graph->constant({1, 1, 4, 16}, keywords::init = inits::from_vector(vState)), nullptr});

  auto aligned = att->apply(state);
  graph->forward();
  ABORT_IF(aligned->shape() != Shape({1, 1, 4, 8}), "aligned shape");
  Values out;
  append(out, aligned);
  return out;
}

// ---- graph_tests.cpp:16-55 (parameter initialisation read-back) -----------
Values graphParamInit() {
  auto graph = newGraph();
  auto pz = graph->param("p_zeros", {2, 3}, keywords::init = inits::zeros);
  auto po = graph->param("p_ones", {2, 3}, keywords::init = inits::ones);
  Values v({1, 2, 3, 4, 5, 6});
  auto pv = graph->param("p_vec", {2, 3}, keywords::init = inits::from_vector(v));
  graph->forward();
  Values out;
  append(out, pz);
  append(out, po);
  append(out, pv);
  return out;
}

}  // namespace

extern "C" int mrn_test_golden(const char* test_case, float* out, size_t capacity, size_t* count) {
  static thread_local std::string err;
  try {
    static const std::map<std::string, Values (*)()> cases = {
        {"operator/dot", opDot},
        {"operator/scalar_mult", opScalarMult},
        {"operator/softmax", opSoftmax},
        {"operator/broadcast", opBroadcast},
        {"operator/transpose", opTranspose},
        {"operator/reductions", opReductions},
        {"operator/concat", opConcat},
        {"operator/layer_norm", opLayerNorm},
        {"rnn/simple", rnnSimple},
        {"rnn/s2s_encoder", rnnS2SEncoder},
        {"attention/context", attentionContext},
        {"graph/param_init", graphParamInit},
    };
    auto it = cases.find(test_case);
    if(it == cases.end())
      return 3;
    Values v = it->second();
    *count = v.size();
    if(out) {
      if(capacity < v.size())
        return 4;
      std::memcpy(out, v.data(), v.size() * sizeof(float));
    }
    return 0;
  } catch(const std::exception& e) {
    fprintf(stderr, "mrn_test_golden(%s): %s\n", test_case, e.what());
    return 1;
  }
}
