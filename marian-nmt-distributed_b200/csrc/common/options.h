// Options: a small typed key/value bag with Marian's accessor API
// (options->get<T>(key), get<T>(key, default), has(key), set(key, value)).
//
// The reference backs this with yaml-cpp (src/common/options.h:29-78) and fills
// it from boost::program_options; the config/CLI layer is out of scope here,
// so options are parsed from a flat "key=value;key=value" string (what the C
// ABI receives) on top of the defaults of src/common/config_parser.cpp:213-560
// that define the benchmark configs.
#pragma once

#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "common/definitions.h"

namespace marian {

class Options {
public:
  Options() {}
  explicit Options(const std::string& spec) { parse(spec); }

  // "a=1;b=foo;dim-vocabs=32000,32000"
  void parse(const std::string& spec) {
    std::stringstream ss(spec);
    std::string item;
    while(std::getline(ss, item, ';')) {
      if(item.empty())
        continue;
      auto eq = item.find('=');
      ABORT_IF(eq == std::string::npos, "Malformed option (expected key=value):", item);
      kv_[trim(item.substr(0, eq))] = trim(item.substr(eq + 1));
    }
  }

  // Takes values of `other` only for keys this bag does not have yet - the
  // semantics of the reference's Options::merge (src/common/options.h:47-51).
  void merge(const Options& other) {
    for(auto& it : other.kv_)
      if(!kv_.count(it.first))
        kv_[it.first] = it.second;
  }
  void merge(Ptr<Options> other) { merge(*other); }
  // Overwrites existing keys.
  void overwrite(const Options& other) {
    for(auto& it : other.kv_)
      kv_[it.first] = it.second;
  }

  bool has(const std::string& key) const { return kv_.count(key) > 0; }

  template <typename T>
  void set(const std::string& key, T value) {
    std::ostringstream os;
    os << value;
    kv_[key] = os.str();
  }
  void set(const std::string& key, const std::string& value) { kv_[key] = value; }
  void set(const std::string& key, const char* value) { kv_[key] = value; }
  void set(const std::string& key, bool value) { kv_[key] = value ? "true" : "false"; }
  void set(const std::string& key, const std::vector<int>& value) {
    std::string s;
    for(size_t i = 0; i < value.size(); ++i)
      s += (i ? "," : "") + std::to_string(value[i]);
    kv_[key] = s;
  }

  template <typename T>
  T get(const std::string& key) const {
    auto it = kv_.find(key);
    ABORT_IF(it == kv_.end(), "Required option has not been set:", key);
    return convert<T>(it->second, key);
  }
  template <typename T>
  T get(const std::string& key, T defaultValue) const {
    auto it = kv_.find(key);
    if(it == kv_.end())
      return defaultValue;
    return convert<T>(it->second, key);
  }

  std::string str() const {
    std::string s;
    for(auto& it : kv_)
      s += it.first + "=" + it.second + ";";
    return s;
  }

  Ptr<Options> clone() const { return Ptr<Options>(new Options(*this)); }

private:
  static std::string trim(const std::string& s) {
    size_t b = s.find_first_not_of(" \t\n");
    size_t e = s.find_last_not_of(" \t\n");
    return b == std::string::npos ? "" : s.substr(b, e - b + 1);
  }

  template <typename T>
  struct tag {};

  template <typename T>
  static T convert(const std::string& v, const std::string& key) {
    return convertImpl(v, key, tag<T>());
  }
  static std::string convertImpl(const std::string& v, const std::string&, tag<std::string>) { return v; }
  static bool convertImpl(const std::string& v, const std::string& key, tag<bool>) {
    if(v == "true" || v == "1" || v == "yes" || v == "on")
      return true;
    if(v == "false" || v == "0" || v == "no" || v == "off" || v.empty())
      return false;
    ABORT("Option is not a boolean:", key, v);
  }
  static int convertImpl(const std::string& v, const std::string&, tag<int>) { return (int)std::stod(v); }
  static size_t convertImpl(const std::string& v, const std::string&, tag<size_t>) { return (size_t)std::stod(v); }
  static float convertImpl(const std::string& v, const std::string&, tag<float>) { return std::stof(v); }
  static double convertImpl(const std::string& v, const std::string&, tag<double>) { return std::stod(v); }
  static std::vector<int> convertImpl(const std::string& v, const std::string&, tag<std::vector<int>>) {
    std::vector<int> out;
    std::stringstream ss(v);
    std::string item;
    while(std::getline(ss, item, ','))
      if(!item.empty())
        out.push_back(std::stoi(item));
    return out;
  }
  static std::vector<std::string> convertImpl(const std::string& v, const std::string&, tag<std::vector<std::string>>) {
    std::vector<std::string> out;
    std::stringstream ss(v);
    std::string item;
    while(std::getline(ss, item, ','))
      if(!item.empty())
        out.push_back(item);
    return out;
  }

  std::map<std::string, std::string> kv_;
};

// Defaults of the training options that shape the hot path
// (reference: src/common/config_parser.cpp:214-467).
inline Ptr<Options> defaultOptions() {
  auto o = New<Options>();
  o->parse(
      "type=s2s;dim-vocabs=0,0;dim-emb=512;dim-rnn=1024;enc-type=bidirectional;enc-cell=gru;enc-cell-depth=1;"
      "enc-depth=1;dec-depth=1;dec-cell=gru;dec-cell-base-depth=2;dec-cell-high-depth=1;skip=false;"
      "layer-normalization=false;right-left=false;tied-embeddings=false;tied-embeddings-src=false;"
      "tied-embeddings-all=false;transformer-heads=8;transformer-dim-ffn=2048;transformer-preprocess=;"
      "transformer-postprocess-emb=d;transformer-postprocess=dan;transformer-dropout=0;"
      "transformer-dropout-attention=0;dropout-rnn=0;dropout-src=0;dropout-trg=0;cost-type=ce-mean;"
      "label-smoothing=0;max-length=50;mini-batch=64;optimizer=adam;learn-rate=0.0001;clip-norm=1;"
      "optimizer-delay=1;exponential-smoothing=0;seed=1234;workspace=2048;sync-sgd=false;inference=false");
  return o;
}

}  // namespace marian
