// Recurrent state containers shared by RNN layers and decoder states.
// reference: src/rnn/types.h:13-98 (State, States)
#pragma once

#include <algorithm>
#include <vector>

#include "graph/expression_graph.h"
#include "graph/expression_operators.h"

namespace marian {
namespace rnn {

struct State {
  Expr output;
  Expr cell;

  // beam-search hypothesis selection (decoding-side; kept for API completeness)
  State select(const std::vector<size_t>& indices, int beamSize) {
    output = atleast_4d(output);
    if(cell)
      cell = atleast_4d(cell);
    int dimDepth = output->shape()[-1];
    int dimTime = output->shape()[-3];
    int dimBatch = (int)indices.size() / beamSize;
    Expr selOut = reshape(rows(flatten_2d(output), indices), {beamSize, dimTime, dimBatch, dimDepth});
    Expr selCell = cell ? reshape(rows(flatten_2d(cell), indices), {beamSize, dimTime, dimBatch, dimDepth}) : nullptr;
    return State{selOut, selCell};
  }
};

class States {
private:
  std::vector<State> states_;

public:
  States() {}
  States(const std::vector<State>& states) : states_(states) {}
  States(size_t num, State state) : states_(num, state) {}

  std::vector<State>::iterator begin() { return states_.begin(); }
  std::vector<State>::iterator end() { return states_.end(); }

  // all time steps stacked on axis -3: [T, B, D]
  Expr outputs() {
    std::vector<Expr> outputs;
    for(auto s : states_)
      outputs.push_back(atleast_3d(s.output));
    if(outputs.size() > 1)
      return concatenate(outputs, keywords::axis = -3);
    return outputs[0];
  }

  State& operator[](size_t i) { return states_[i]; }
  const State& operator[](size_t i) const { return states_[i]; }
  State& back() { return states_.back(); }
  const State& back() const { return states_.back(); }
  State& front() { return states_.front(); }
  const State& front() const { return states_.front(); }
  size_t size() const { return states_.size(); }
  void push_back(const State& state) { states_.push_back(state); }

  States select(const std::vector<size_t>& indices, int beamSize) {
    States selected;
    for(auto& state : states_)
      selected.push_back(state.select(indices, beamSize));
    return selected;
  }

  void reverse() { std::reverse(states_.begin(), states_.end()); }
  void clear() { states_.clear(); }
};

}  // namespace rnn
}  // namespace marian
