// Model type -> encoder/decoder composition.
// reference: src/models/model_factory.cpp:61-192 (only the compositions the
// benchmark configs use: "transformer" and "s2s").
#pragma once

#include "models/encdec.h"

namespace marian {
namespace models {

Ptr<EncoderDecoder> from_options(Ptr<Options> options);

}  // namespace models
}  // namespace marian
