// TEST INFRASTRUCTURE (oracle/Makefile target refmodels): the reference's umbrella header, redirected to this
// repo's host engine, so that /root/reference/src/models/{transformer,s2s}.h compile UNCHANGED against it.
#pragma once
#include "graph/expression_graph.h"
#include "graph/expression_operators.h"
#include "layers/generic.h"
#include "models/encdec.h"
#include "rnn/rnn.h"
