#!/bin/bash
# Build container only (needs /root/reference). Compiles (syntax only) every example driver of the reference, UNCHANGED,
#   (a) against the reference's own headers (oracle/shim stands in for Blitz++ / GSL / HDF5)  and
#   (b) against the host mirror weldform_b200/host/,
# and prints which do. Usage: scripts/sweep_examples.sh [out.txt]
REPO=$(cd "$(dirname "$0")/.." && pwd); REF=${REF:-/root/reference}; OUT=${1:-/dev/stdout}
T=$(mktemp -d)
ls $REF/examples/*.cpp | xargs -P 8 -I{} bash -c '
  f={}; b=$(basename $f .cpp)
  a=FAIL; m=FAIL
  g++ -std=c++14 -fsyntax-only -fopenmp -fpermissive -w -include iomanip -D_USE_MATH_DEFINES -I'$REPO'/oracle/shim -I'$REPO'/oracle -I'$REF'/Source -I'$REF'/External $f > '$T'/$b.ref.log 2>&1 && a=ok
  g++ -std=c++14 -fsyntax-only -fopenmp -fpermissive -w -D_USE_MATH_DEFINES -I'$REPO'/weldform_b200/host -I'$REPO'/include $f > '$T'/$b.mir.log 2>&1 && m=ok
  why=""; [ $m = FAIL ] && why=$(grep -m1 error '$T'/$b.mir.log | sed "s/.*error: //" | cut -c1-110)
  printf "%-44s reference:%-4s mirror:%-4s %s\n" $b $a $m "$why"' | sort > $T/all.txt
{
  echo "examples: $(wc -l < $T/all.txt); compile against the reference's headers: $(grep -c 'reference:ok' $T/all.txt); against the mirror: $(grep -c 'mirror:ok' $T/all.txt); both: $(grep 'reference:ok' $T/all.txt | grep -c 'mirror:ok')"
  cat $T/all.txt
} > $OUT
rm -rf $T
