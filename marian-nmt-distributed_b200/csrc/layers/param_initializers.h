// Parameter / constant initialisers.
//
// Reproduces the reference's value streams (src/layers/param_initializers.h:27-33,
// .cu:68-71): glorot_uniform draws uniform(+-sqrt(6/(d0+d1))) from a NEW
// std::default_random_engine(Config::seed++) per tensor, values are generated
// on the host and uploaded.  Because the seed is consumed when the tensor is
// first initialised (inside forward(), tape order), parameter creation order
// defines the stream - models here create parameters in the reference's order.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
#include <random>
#include <vector>

#include "common/definitions.h"
#include "tensors/tensor.h"

namespace marian {

// The reference keeps the global seed in Config::seed (src/common/config.h).
struct Config {
  static size_t seed;
};

namespace inits {

inline void zeros(Tensor t) {
  t->set(0.f);
}
inline void ones(Tensor t) {
  t->set(1.f);
}
inline std::function<void(Tensor)> from_value(float v) {
  return [v](Tensor t) { t->set(v); };
}

template <class Distribution>
void distribution(std::vector<float>& vals, float a, float b) {
  std::default_random_engine engine(Config::seed++);
  Distribution dist(a, b);
  auto gen = std::bind(dist, engine);
  std::generate(vals.begin(), vals.end(), gen);
}

template <class Distribution>
void distribution(Tensor t, float a, float b) {
  std::vector<float> vals(t->size());
  distribution<Distribution>(vals, a, b);
  t->set(vals);
}

inline std::function<void(Tensor)> normal(float scale = 0.1f) {
  return [scale](Tensor t) { distribution<std::normal_distribution<float>>(t, 0, scale); };
}
inline std::function<void(Tensor)> uniform(float scale = 0.1f) {
  return [scale](Tensor t) { distribution<std::uniform_real_distribution<float>>(t, -scale, scale); };
}
inline void glorot_uniform(Tensor t) {
  float scale = sqrtf(6.0f / (t->shape()[0] + t->shape()[1]));
  distribution<std::uniform_real_distribution<float>>(t, -scale, scale);
}
inline void glorot_normal(Tensor t) {
  float scale = sqrtf(2.0f / (t->shape()[0] + t->shape()[1]));
  distribution<std::normal_distribution<float>>(t, 0, scale);
}

inline std::function<void(Tensor)> from_vector(const std::vector<float>& v) {
  return [v](Tensor t) { t->set(v); };
}
inline std::function<void(Tensor)> from_vector(const std::vector<size_t>& v) {
  std::vector<float> vf(v.begin(), v.end());
  return from_vector(vf);
}

}  // namespace inits
}  // namespace marian
