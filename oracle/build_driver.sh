#!/bin/bash
# TEST INFRASTRUCTURE (build container only). Builds ONE example driver of the reference, unmodified, into oracle/_ref/drv_<name>:
#     oracle/build_driver.sh <name> <Source dir: the reference's or the patched scratch copy, see Makefile> [REF=/root/reference]
# The reference defines its solver loops `inline` in SolverFraser.cpp / SolverKickDrift.cpp / SolverLeapfrog.cpp, outside Domain.h (its cmake
# build lists them with the driver); an example that calls a loop without including its file gets that file appended to the SAME translation
# unit here (two #include lines on stdin). No reference source is copied or changed.
set -e
name=$1; srcdir=$2; REF=${REF:-/root/reference}
here=$(cd "$(dirname "$0")" && pwd); cd "$here"; mkdir -p _ref
src=$REF/examples/$name.cpp
code=$(grep -v '^\s*//' "$src")
tu="#include \"$src\"\n"
for pair in "SolveDiffUpdateKickDrift:SolverKickDrift.cpp" "SolveDiffUpdateLeapFrog:SolverLeapfrog.cpp" "SolveDiffUpdateFraser:SolverFraser.cpp"; do
  fn=${pair%%:*}; file=${pair##*:}
  if echo "$code" | grep -q "$fn" && ! grep -q "^\s*#include\s*\"$file\"" "$src"; then tu="$tu#include \"$file\"\n"; fi
done
CXX=g++; [ -x /usr/bin/g++ ] && CXX=/usr/bin/g++
printf "$tu" | $CXX -x c++ -std=c++14 -O3 -fopenmp -fpermissive -w -ffp-contract=off -include iomanip -D_USE_MATH_DEFINES -Ishim -I. -I"$srcdir" -I$REF/External - \
    -x none $REF/External/LSDynaReader/src/lsdynaReader.cpp -o _ref/drv_$name
