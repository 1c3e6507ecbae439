// Shim for the one Boost header the reference's kernel translation units pull in
// (src/common/keywords.h:24 -> boost/any.hpp): Boost is not installed here.
#pragma once
#include <any>
namespace boost {
using any = std::any;
using std::any_cast;
using bad_any_cast = std::bad_any_cast;
}  // namespace boost
