#!/bin/bash
# Which of the reference's host-side headers compile UNCHANGED against this repo's engine (csrc/)?
#   models/transformer.h, models/s2s.h : yes - built into oracle/_ref/libmarian_oracle_refmodels.so and run
#                                        (make -C oracle refmodels; tests/test_reference_models.py)
#   rnn/{rnn,cells,attention,constructors}.h : attempted here with g++ -fsyntax-only; the errors are recorded.
# Needs /root/reference (build container only).  Output: profiles/boundary_check_r02.txt
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=/root/reference/src
OUT=$ROOT/profiles/boundary_check_r02.txt
TMP=$(mktemp -d)
INC="-I$ROOT/oracle/cpu -I$ROOT/marian-nmt-distributed_b200/csrc -I$ROOT/include"
{
echo "# boundary check: reference headers compiled where they lie against csrc/ (g++ -std=c++17 -fsyntax-only, CPU operator layer)"
for h in models/transformer.h models/s2s.h; do
  printf '#include "marian.h"\n#include "%s/%s"\nint main(){return 0;}\n' $REF $h > $TMP/tu.cpp
  n=$(g++ -std=c++17 -DMRN_ORACLE_CPU=1 -I$ROOT/oracle/ref_shims $INC -fsyntax-only $TMP/tu.cpp 2>&1 | grep -c "error")
  echo "$h (umbrella marian.h -> oracle/ref_shims/marian.h): $n errors"
done
# the reference's rnn layer on top of this repo's graph / layer API
mkdir -p $TMP/shim/rnn $TMP/shim/layers
for f in rnn.h cells.h attention.h types.h constructors.h; do ln -sf $REF/rnn/$f $TMP/shim/rnn/$f; done
ln -sf $REF/layers/factory.h $TMP/shim/layers/factory.h
printf '#pragma once\n#include <cassert>\n#include "graph/expression_graph.h"\n#include "graph/expression_operators.h"\n' > $TMP/shim/marian.h
printf '#include "marian.h"\n#include "rnn/rnn.h"\n#include "rnn/constructors.h"\nint main(){return 0;}\n' > $TMP/tu.cpp
echo
echo "rnn/rnn.h + rnn/cells.h + rnn/attention.h + rnn/constructors.h (+ layers/factory.h) of the reference: errors (first line of each)"
g++ -std=c++17 -DMRN_ORACLE_CPU=1 -I$TMP/shim $INC -fsyntax-only $TMP/tu.cpp 2>&1 | grep "error" | sed "s#$TMP/shim/#reference:src/#; s#$ROOT/##" | cut -c1-220 | head -40
} > $OUT
rm -rf $TMP
cat $OUT
