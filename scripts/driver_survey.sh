#!/bin/bash
# Which of the reference's example drivers compile UNCHANGED against the host mirror (weldform_b200/host/)?  Build container only.
#   usage: scripts/driver_survey.sh [-v]    prints "ok=<n> fail=<m>", with -v the first error of every failing driver
cd "$(dirname "$0")/.." || exit 1
ok=0; fail=0
for f in /root/reference/examples/*.cpp; do
  if g++ -std=c++14 -fsyntax-only -fpermissive -w -fopenmp -D_USE_MATH_DEFINES -Iweldform_b200/host -Iinclude "$f" 2>/tmp/wf_survey_err.txt; then
    ok=$((ok+1)); [ "$1" = "-l" ] && echo "OK $(basename "$f")"
  else
    fail=$((fail+1)); [ "$1" = "-v" ] && echo "== $(basename "$f"): $(grep -m1 'error' /tmp/wf_survey_err.txt | cut -c1-200)"
  fi
done
echo "ok=$ok fail=$fail"
