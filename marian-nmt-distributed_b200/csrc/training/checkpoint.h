// Checkpoint / resume in the reference's wire format.
//
// Reference: ExpressionGraph::save / load (src/graph/expression_graph.h:442-502): every parameter as
// a float32 array under its Marian name (namespace prefix stripped) in one .npz, written in the order of
// Parameters::getMap() (a std::map: sorted by name); EncoderDecoder::save (src/models/encdec.h:201-229,
// 275-290) appends the model description as the char array "special:model.yml" (the keys of
// modelFeatures_ + version).  load() creates the parameters from the file BEFORE the model is built;
// the model code then finds them by name (ExpressionGraph::param returns existing parameters).
// Added here (the reference cannot resume its optimizer): the Adam moments and step counter in
// "<path>.optimizer.npz", stored per parameter name so that they survive the different arena order
// of a reloaded model (creation order = file order).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "common/npz.h"
#include "common/options.h"
#include "graph/expression_graph.h"
#include "layers/param_initializers.h"
#include "optimizers/optimizers.h"

namespace marian {
namespace checkpoint {

// keys of EncoderDecoder::modelFeatures_ (src/models/encdec.h:244-272), in that order
inline const std::vector<std::string>& modelFeatures() {
  static const std::vector<std::string> keys = {
      "type", "dim-vocabs", "dim-emb", "dim-rnn", "enc-cell", "enc-type", "enc-cell-depth", "enc-depth", "dec-depth", "dec-cell",
      "dec-cell-base-depth", "dec-cell-high-depth", "skip", "layer-normalization", "right-left", "special-vocab", "tied-embeddings",
      "tied-embeddings-src", "tied-embeddings-all", "transformer-heads", "transformer-dim-ffn", "transformer-preprocess",
      "transformer-postprocess", "transformer-postprocess-emb"};
  return keys;
}

// block-style YAML as yaml-cpp's default emitter writes it: "key: value", sequences as "  - item"
inline std::string modelYaml(Ptr<Options> options) {
  std::string y;
  for(auto& key : modelFeatures()) {
    if(!options->has(key))
      continue;
    std::string v = options->get<std::string>(key);
    if(key == "dim-vocabs" || key == "special-vocab") {
      y += key + ":\n";
      for(auto& item : options->get<std::vector<std::string>>(key))
        y += "  - " + item + "\n";
    } else {
      y += key + ": " + (v.empty() ? "\"\"" : v) + "\n";
    }
  }
  y += "version: v1.2.1+b200\n";
  return y;
}

// "key: value" / "key:\n  - a\n  - b" back into an option string (what a reloaded model overrides)
inline Options parseModelYaml(const std::string& yaml) {
  Options o;
  std::stringstream ss(yaml);
  std::string line, listKey, listVal;
  auto flush = [&] {
    if(!listKey.empty())
      o.set(listKey, listVal);
    listKey.clear();
    listVal.clear();
  };
  while(std::getline(ss, line)) {
    if(line.empty())
      continue;
    if(line.rfind("  - ", 0) == 0 || line.rfind("- ", 0) == 0) {
      std::string item = line.substr(line.find("- ") + 2);
      listVal += (listVal.empty() ? "" : ",") + item;
      continue;
    }
    flush();
    auto c = line.find(':');
    if(c == std::string::npos)
      continue;
    std::string key = line.substr(0, c), val = c + 1 < line.size() ? line.substr(c + 1) : "";
    size_t b = val.find_first_not_of(' ');
    val = b == std::string::npos ? "" : val.substr(b);
    if(val == "\"\"" || val == "''")
      val.clear();
    if(val.empty() && ss.peek() == ' ')
      listKey = key;
    else
      o.set(key, val);
  }
  flush();
  return o;
}

inline npz::Array floatArray(const std::vector<int>& shape, const std::vector<float>& v) {
  npz::Array a;
  a.shape = shape;
  a.kind = 'f';
  a.bytes.assign((const char*)v.data(), (const char*)v.data() + v.size() * sizeof(float));
  return a;
}
inline npz::Array textArray(const std::string& s) {
  npz::Array a;
  a.shape = {(int)s.size() + 1};  // the reference stores the terminating 0 as well (config.cpp:65)
  a.kind = 'i';
  a.bytes.assign(s.begin(), s.end());
  a.bytes.push_back('\0');
  return a;
}

inline std::map<std::string, Expr> paramsByName(Ptr<ExpressionGraph> graph) {
  std::map<std::string, Expr> m;
  for(auto p : *graph->params())
    m[p->name()] = p;
  return m;
}

inline void saveModel(Ptr<ExpressionGraph> graph, Ptr<Options> options, const std::string& path) {
  std::vector<std::pair<std::string, npz::Array>> out;
  for(auto& kv : paramsByName(graph)) {
    std::vector<float> v;
    kv.second->val()->get(v);
    std::vector<int> shape;
    for(auto d : kv.second->shape())
      shape.push_back(d);
    out.push_back({kv.first, floatArray(shape, v)});
  }
  out.push_back({"special:model.yml", textArray(modelYaml(options))});
  npz::save(path, out);
}

// Creates the parameters of `path` in `graph` (before the model is built) and returns the model options
// stored with them (empty if the file has no special:model.yml).
inline Options loadModel(Ptr<ExpressionGraph> graph, const std::string& path) {
  using namespace keywords;
  Options stored;
  for(auto& kv : npz::load(path)) {
    if(kv.first.rfind("special:", 0) == 0) {  // reference: skipped by ExpressionGraph::load
      if(kv.first == "special:model.yml")
        stored = parseModelYaml(kv.second.text());
      continue;
    }
    ABORT_IF(kv.second.kind != 'f', "checkpoint: parameter is not float32:", kv.first);
    Shape shape;
    if(kv.second.shape.size() == 1) {  // reference :458-462: vectors become [1, n]
      shape = Shape{1, kv.second.shape[0]};
    } else {
      shape.resize(kv.second.shape.size());
      for(size_t i = 0; i < kv.second.shape.size(); ++i)
        shape.set((int)i, kv.second.shape[i]);
    }
    std::vector<float> v(kv.second.floats(), kv.second.floats() + kv.second.elements());
    graph->param(kv.first, shape, init = inits::from_vector(v));
  }
  return stored;
}

// ---- optimizer state (this repo's addition) ----
inline void saveAdam(Ptr<ExpressionGraph> graph, Ptr<Adam> adam, const std::string& path) {
  ABORT_IF(!adam->mt(), "checkpoint: the optimizer has not made a step yet");
  std::vector<float> mt, vt;
  adam->mt()->get(mt);
  adam->vt()->get(vt);
  const float* base = graph->params()->vals()->data();
  std::vector<std::pair<std::string, npz::Array>> out;
  for(auto& kv : paramsByName(graph)) {
    size_t off = (size_t)(kv.second->val()->data() - base), n = kv.second->val()->size();
    std::vector<int> shape;
    for(auto d : kv.second->shape())
      shape.push_back(d);
    out.push_back({"adam_mt:" + kv.first, floatArray(shape, std::vector<float>(mt.begin() + off, mt.begin() + off + n))});
    out.push_back({"adam_vt:" + kv.first, floatArray(shape, std::vector<float>(vt.begin() + off, vt.begin() + off + n))});
  }
  out.push_back({"special:optimizer.yml", textArray("type: adam\nsteps: " + std::to_string(adam->steps()) + "\n")});
  npz::save(path, out);
}

// The moments are scattered into the flat state when the optimizer allocates it (first update after
// the reload): by then the reloaded parameters have their arena offsets.
inline void loadAdam(Ptr<ExpressionGraph> graph, Ptr<Adam> adam, const std::string& path) {
  auto arrays = std::make_shared<std::map<std::string, npz::Array>>();
  size_t steps = 0;
  for(auto& kv : npz::load(path)) {
    if(kv.first == "special:optimizer.yml")
      steps = parseModelYaml(kv.second.text()).get<size_t>("steps", 0);
    else
      (*arrays)[kv.first] = kv.second;
  }
  Weak<ExpressionGraph> weak = graph;
  adam->restoreOnAllocation(steps, [arrays, weak](Tensor mt, Tensor vt) {
    auto g = weak.lock();
    ABORT_IF(!g, "checkpoint: graph is gone");
    const float* base = g->params()->vals()->data();
    std::vector<float> m(mt->size(), 0.f), v(vt->size(), 0.f);
    for(auto p : *g->params()) {
      size_t off = (size_t)(p->val()->data() - base), n = p->val()->size();
      auto im = arrays->find("adam_mt:" + p->name()), iv = arrays->find("adam_vt:" + p->name());
      ABORT_IF(im == arrays->end() || iv == arrays->end(), "checkpoint: no optimizer state for parameter", p->name());
      ABORT_IF(im->second.elements() != n || iv->second.elements() != n, "checkpoint: optimizer state has the wrong size for", p->name());
      std::copy(im->second.floats(), im->second.floats() + n, m.begin() + off);
      std::copy(iv->second.floats(), iv->second.floats() + n, v.begin() + off);
    }
    mt->set(m);
    vt->set(v);
  });
}

}  // namespace checkpoint
}  // namespace marian
