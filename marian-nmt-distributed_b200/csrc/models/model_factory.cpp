// reference: src/models/model_factory.cpp:61-192
#include "models/model_factory.h"
#include "models/s2s.h"
#include "models/transformer.h"

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
