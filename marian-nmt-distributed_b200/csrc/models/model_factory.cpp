// reference: src/models/model_factory.cpp:61-192
#include "models/model_factory.h"
// MRN_REFERENCE_MODELS (oracle/Makefile, target refmodels - TEST INFRASTRUCTURE): the encoder / decoder classes come
// from the REFERENCE's own src/models/{transformer,s2s}.h, compiled where they lie against this repo's graph /
// operator / layer / rnn API (their "marian.h" umbrella is redirected by oracle/ref_shims/marian.h).  That is the
// north-star's "models compile unchanged" boundary as a build target; tests/test_reference_models.py runs both
// model codes on the same batches and compares cost, logits and every gradient.
#ifdef MRN_REFERENCE_MODELS
#include MRN_REFERENCE_S2S_H
#include MRN_REFERENCE_TRANSFORMER_H
#else
#include "models/s2s.h"
#include "models/transformer.h"
#endif

namespace marian {
namespace models {

Ptr<EncoderDecoder> from_options(Ptr<Options> options) {
  std::string type = options->get<std::string>("type");
  auto encdec = New<EncoderDecoder>(options);

  auto sub = [&](const std::string& prefix, size_t index) {
    auto o = options->clone();
    o->set("prefix", prefix);
    o->set("index", index);
    return o;
  };

  if(type == "transformer") {
    encdec->push_back(Ptr<EncoderBase>(New<EncoderTransformer>(sub("encoder", 0))));
    encdec->push_back(Ptr<DecoderBase>(New<DecoderTransformer>(sub("decoder", 1))));
  } else if(type == "s2s") {
    encdec->push_back(Ptr<EncoderBase>(New<EncoderS2S>(sub("encoder", 0))));
    encdec->push_back(Ptr<DecoderBase>(New<DecoderS2S>(sub("decoder", 1))));
  } else {
    ABORT("Unknown model type:", type);
  }
  return encdec;
}

}  // namespace models
}  // namespace marian
